import sys,os,ctypes as C
sys.path.insert(0,'tests')
import numpy as np, orc, parity
pkg=orc._load_package()
cfgs=[("cfg2_lcurve48",48,8e-3,40,"lcurve",{},2048),("cfg3_lcurve56",56,7e-3,40,"lcurve",{},2048),("cfg4_gcv",48,8e-3,60,"gcv",{},512),("cfg4_chi2",48,8e-3,60,"chi2",{"Chi2Factor":1.02},1024),("cfg5_mdp",32,10e-3,60,"mdp",{"NoiseLevel":1e-3},1024)]
for name,nTE,TE,nT2,Reg,extra,nvox in cfgs:
    img=orc.mock_image(nvox,nTE,TE,seed=7)
    o=orc.make_t2map_opts((nvox,1,1),nTE,nT2,TE,Reg=Reg,ngpus=1,**extra); p=orc.make_t2part_opts((nvox,1,1),nT2)
    ref,_=orc.t2map(img,o,p)
    arrs,out=orc.alloc_outputs(nvox,nTE,nT2,part=True)
    rc=pkg.lib().decaes_t2map(img.ctypes.data,C.byref(o),C.byref(p),C.byref(out)); assert rc==0
    arrs["dist"]=arrs["dist"].reshape(nT2,nvox).T
    r=parity.compare(ref,arrs)
    print(name,"refine=",os.environ.get("DECAES_REFINE","default"),"out=%d flips=%d (%.1f%%) med_dlog=%.2e same_mu_out=%d support_diff=%d maxrel_same_support=%.1e"%(r['voxels_out_of_tolerance'],r['mu_flips'],100*r['mu_flip_frac'],r['mu_flip_median_dlog'],r['out_of_tolerance_same_mu'],r['support_diff'],r['dist_max_rel_same_support']))
