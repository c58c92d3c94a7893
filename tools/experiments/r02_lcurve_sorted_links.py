p='decaes.jl_b200/csrc/voxel.cuh'
s=open(p).read()
def rep(old,new,cnt=1):
    global s
    assert s.count(old)==cnt, (s.count(old), old[:70])
    s=s.replace(old,new)
# smem layout: neighbour links
rep("  int A, b, u, x, w, bd, sig, fit, slot_mu, slot_lmu, slot_r2, slot_x2, slot_mask, idx, bar, total_bytes;",
    "  int A, b, u, x, w, bd, sig, fit, slot_mu, slot_lmu, slot_r2, slot_x2, slot_mask, idx, bar, lc_nb, total_bytes;")
rep("    bar = o, o += 2;\n","    bar = o, o += 2;\n    lc_nb = o, o += gram ? 2 * DECAES_LC_MAX / 8 : 0;  // left / right neighbour of every L-curve point (one byte each)\n")
rep("  uint64_t *bar;\n","  uint64_t *bar;\n  unsigned char *lc_nb;             // sorted-order links of the L-curve point cache: [2 i] left, [2 i + 1] right neighbour (0xff: none)\n  unsigned long long lc_dirty;      // points whose neighbours changed since their curvature was last computed\n")
rep("    bar = (uint64_t *)(smem + L.bar);\n","    bar = (uint64_t *)(smem + L.bar);\n    lc_nb = (unsigned char *)(smem + L.lc_nb);\n    lc_dirty = 0ull;\n")
# lc_eval: link the new point
rep("""    i = npts;
    if (npts < DECAES_LC_MAX) {
      if (lane == 0) pts[4 * i] = t, pts[4 * i + 1] = xi, pts[4 * i + 2] = eta, pts[4 * i + 3] = -CUDART_INF;
      npts++;
    } else {
      i = DECAES_LC_MAX - 1;
      n_overflow++;
    }
    __syncwarp();
    return i;
  }
""","""    i = npts;
    if (npts < DECAES_LC_MAX) {
      if constexpr (GRAM) {
        // sorted-order links: the new point goes between its nearest cached neighbours, whose curvatures (and its own)
        // are the only ones that change
        int im, ip;
        lc_neighbours(t, npts, im, ip);
        unsigned char *nb = lc_nb;
        SH(nb);
        if (lane == 0) {
          nb[2 * i] = (unsigned char)(im == 0x7fffffff ? 0xff : im), nb[2 * i + 1] = (unsigned char)(ip == 0x7fffffff ? 0xff : ip);
          if (im != 0x7fffffff) nb[2 * im + 1] = (unsigned char)i;
          if (ip != 0x7fffffff) nb[2 * ip] = (unsigned char)i;
        }
        lc_dirty |= 1ull << i;
        if (im != 0x7fffffff) lc_dirty |= 1ull << im;
        if (ip != 0x7fffffff) lc_dirty |= 1ull << ip;
      }
      if (lane == 0) pts[4 * i] = t, pts[4 * i + 1] = xi, pts[4 * i + 2] = eta, pts[4 * i + 3] = -CUDART_INF;
      npts++;
    } else {
      i = DECAES_LC_MAX - 1;
      n_overflow++;
      lc_dirty = ~0ull;
    }
    __syncwarp();
    return i;
  }

  // nearest cached abscissae on either side of x (src/lsqnonneg.jl:954-959), first index on ties; 0x7fffffff = none
  __device__ __noinline__ void lc_neighbours(double x, int npts, int &im, int &ip) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    const double *pts = lc_pts_p;
    unsigned long long km = 0ull, kp = ~0ull, best;
    im = 0x7fffffff, ip = 0x7fffffff;
    _Pragma("unroll 1") for (int k = lane; k < npts; k += 32) {
      const double _x = pts[4 * k];
      const unsigned long long kk = dkey(_x);
      if (_x < x && kk > km) km = kk, im = k;
      if (x < _x && kk < kp) kp = kk, ip = k;
    }
    im = warp_argmax_bits(km, im, best);
    if (best == 0ull) im = 0x7fffffff;
    ip = warp_argmin_bits(kp, ip, best);
    if (best == ~0ull) ip = 0x7fffffff;
  }
""")
# update_curvature
a=s.index("  __device__ __noinline__ void lc_update_curvature(const double *sx, const int *si, int npts, double tlx, double tly, double brx,")
b=s.index("  // mapfindmax over the curvatures: first maximum under Base.isless (NaN is maximal)")
new='''  // update_curvature!  src/lsqnonneg.jl:948-972 for the four points of the current state.  The reference recomputes all
  // four at every step from the nearest cached abscissae on either side; the result only changes when those neighbours
  // do, so a point is recomputed when it is new or a new point was linked next to it (lc_dirty) - same values, about a
  // third of the work.  A state abscissa that is only isapprox-equal to its cached key (the reference then finds the
  // point among its own neighbours) takes the reference's scan.
  __device__ __noinline__ void lc_update_curvature(const double *sx, const int *si, int npts, double tlx, double tly, double brx,
                                      double bry, double Ctol) {
    const int lane = this->lane;
    VIEWG(double, lc_pts_p);
    double *pts = lc_pts_p;
    const unsigned char *nb = lc_nb;
    SHG(nb);
    _Pragma("unroll 1") for (int q = 0; q < 4; q++) {
      const int pi = si[q];
      const double x = sx[q];
      bool exact = false;
      if constexpr (GRAM) {
        exact = (x == pts[4 * pi]);
        if (exact && !((lc_dirty >> pi) & 1ull)) continue;
      }
      const double px = pts[4 * pi + 1], py = pts[4 * pi + 2];
      double C = -CUDART_INF;
      if (fmin(norm2(px, py, tlx, tly), norm2(px, py, brx, bry)) > Ctol) {
        int im, ip;
        if (exact) {
          im = nb[2 * pi], ip = nb[2 * pi + 1];
          if (im == 0xff) im = 0x7fffffff;
          if (ip == 0xff) ip = 0x7fffffff;
        } else {
          lc_neighbours(x, npts, im, ip);
        }
        double mx = px, my = py, qx = px, qy = py;
        if (im != 0x7fffffff) mx = pts[4 * im + 1], my = pts[4 * im + 2];
        if (ip != 0x7fffffff) qx = pts[4 * ip + 1], qy = pts[4 * ip + 2];
        C = menger(mx, my, px, py, qx, qy);
      }
      __syncwarp();
      if (lane == 0) pts[4 * pi + 3] = C;
      __syncwarp();
      if (exact) lc_dirty &= ~(1ull << pi);
      else lc_dirty |= 1ull << pi;
    }
  }

'''
s=s[:a]+new+s[b:]
# reset at the start of the search
rep("    int npts = 0, nst = 0;\n    double sx[4];","    int npts = 0, nst = 0;\n    lc_dirty = 0ull;\n    double sx[4];")
open(p,'w').write(s)
print("ok")
