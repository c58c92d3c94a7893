"""Command line interface: a Python mirror of `decaes` / `DECAES.main` (src/main.jl) over the GPU library.

    python -m decaes_cli image.nii.gz --T2map --T2part --TE 7e-3 --nT2 40 --T2Range 10e-3 2.0 \
           --SPWin 10e-3 25e-3 --MPWin 25e-3 200e-3 --Reg lcurve --output results/
    python -m decaes_cli @settings.txt            (one argument per line, like the reference's settings files)

Same flags, same interdependencies and messages, same output files (<name>.t2dist.mat, <name>.t2maps.mat,
<name>.t2parts.mat).  Differences: `--bet` (FSL brain extraction) is refused, PAR/REC inputs
are not read, MAT files are written as v5 (fileio.py), and `--ngpus` selects how many devices share a volume.
When both --T2map and --T2part are given the T2part maps come from the fused epilogue of the same kernel.
"""
import argparse
import os
import sys
import time
import warnings

import numpy as np

from . import T2mapOptions, T2mapSEcorr, T2partOptions, T2partSEcorr
from .fileio import (ALLOWED_FILE_SUFFIXES_STRING, chop_allowed_suffix, is_allowed_suffix, load_image, save_mat)

T2MAP_FIELDS = {"TE", "nT2", "T2Range", "Reg", "T1", "Threshold", "MinRefAngle", "nRefAngles", "nRefAnglesMin", "Chi2Factor",
                "NoiseLevel", "RefConAngle", "SetFlipAngle", "SaveResidualNorm", "SaveDecayCurve", "SaveRegParam",
                "SaveNNLSBasis", "legacy"}
T2PART_FIELDS = {"nT2", "T2Range", "SPWin", "MPWin", "Sigmoid", "legacy"}


def build_parser():
    p = argparse.ArgumentParser(prog="decaes", fromfile_prefix_chars="@", allow_abbrev=False,
                                description="DECAES on B200: T2 distributions (T2mapSEcorr) and T2 parts (T2partSEcorr)")
    p.add_argument("input", nargs="*", help=f"one or more input filenames. Valid file types are limited to: {ALLOWED_FILE_SUFFIXES_STRING}")
    p.add_argument("--mask", "-m", nargs="+", default=[], help="one or more mask filenames, applied by elementwise multiplication")
    p.add_argument("--output", "-o", nargs="+", default=[], help="one or more output directories")
    p.add_argument("--T2map", action="store_true", help="compute T2 distributions from 4D multi spin-echo images")
    p.add_argument("--T2part", action="store_true", help="analyse 4D T2 distributions to produce parameter maps")
    p.add_argument("--quiet", "-q", action="store_true", help="suppress printing to the terminal")
    p.add_argument("--dry", action="store_true", help="execute dry run of processing without saving any results")
    p.add_argument("--legacy", action="store_const", const=True, default=None, help="(deprecated) use legacy settings and algorithms from the original MATLAB pipeline")
    g = p.add_argument_group("T2map/T2part required parameters")
    g.add_argument("--MatrixSize", nargs=3, type=int, help="inferred from the input image")
    g.add_argument("--nTE", type=int, help="inferred from the input image")
    g.add_argument("--TE", type=float, help="inter-echo spacing (seconds). Required when --T2map is passed")
    g.add_argument("--nT2", type=int, help="number of T2 components. Required when --T2map is passed")
    g.add_argument("--T2Range", nargs=2, type=float, help="minimum and maximum T2 values (seconds)")
    g.add_argument("--SPWin", nargs=2, type=float, help="short peak window (seconds). Required when --T2part is passed")
    g.add_argument("--MPWin", nargs=2, type=float, help="middle peak window (seconds). Required when --T2part is passed")
    g.add_argument("--Reg", type=str, help='one of "lcurve", "gcv", "chi2", "mdp", or "none"')
    g.add_argument("--RegParams", nargs="+", type=float, default=[], help='required if --Reg="chi2" or --Reg="mdp"')
    g.add_argument("--Chi2Factor", type=float, help="(deprecated) use --RegParams instead")
    o = p.add_argument_group("T2map/T2part optional parameters")
    o.add_argument("--T1", type=float)
    o.add_argument("--Sigmoid", type=float)
    o.add_argument("--Threshold", type=float)
    o.add_argument("--B1map", nargs="+", default=[], help="one or more B1 map filenames (flip angles in degrees)")
    o.add_argument("--nRefAngles", type=int)
    o.add_argument("--nRefAnglesMin", type=int)
    o.add_argument("--MinRefAngle", type=float)
    o.add_argument("--SetFlipAngle", type=float)
    o.add_argument("--RefConAngle", type=float)
    o.add_argument("--SaveDecayCurve", action="store_true")
    o.add_argument("--SaveNNLSBasis", action="store_true")
    o.add_argument("--SaveRegParam", action="store_true")
    o.add_argument("--SaveResidualNorm", action="store_true")
    o.add_argument("--bet", action="store_true", help="FSL BET masks: not available here, pass --mask instead")
    o.add_argument("--ngpus", type=int, default=0, help="extension: number of GPUs sharing each volume (0 = all visible)")
    return p


def parse_cli(args):
    """parse_cli + handle_cli_deprecations! + verify_cli_args! + clean_cli_args!  (src/main.jl:424-484)."""
    opts = vars(build_parser().parse_args(args))
    if opts.get("legacy") is not None:  # warn_deprecated_future_removed(:legacy)  src/main.jl:441-443, 457
        warnings.warn("The flag --legacy is deprecated and will be removed in future releases.")
    if opts.get("Chi2Factor") is not None:
        if opts["RegParams"]:
            raise SystemExit("The flag --RegParams and the deprecated flag --Chi2Factor were both passed; use --RegParams only.")
        warnings.warn("The flag --Chi2Factor is deprecated and will be removed in future releases; use --RegParams instead.")
        opts["RegParams"] = [opts["Chi2Factor"]]
    opts.pop("Chi2Factor", None)
    if not (opts["T2map"] or opts["T2part"]):
        raise SystemExit("At least one of --T2map or --T2part must be passed")
    if opts["bet"]:
        raise SystemExit("--bet needs the FSL toolbox; create the mask beforehand and pass it with --mask")
    if opts["Reg"] == "chi2":
        assert len(opts["RegParams"]) == 1, 'Must set chi2 factor via --RegParams when --Reg="chi2"'
        opts["Chi2Factor"] = opts["RegParams"][0]
    elif opts["Reg"] == "mdp":
        assert len(opts["RegParams"]) == 1, 'Must set noise level via --RegParams when --Reg="mdp"'
        opts["NoiseLevel"] = opts["RegParams"][0]
    return opts


def get_file_infos(opts):
    """src/main.jl:511-575: pair every input with its output folder, mask and B1 map."""
    inputs = opts["input"]
    assert inputs, "At least one input file is required"
    inputfiles = [f for f in inputs if is_allowed_suffix(f)]
    if not inputfiles:
        if inputs and os.path.isfile(inputs[0]):
            raise SystemExit("No valid file types were found for processing, but a file name was passed.\n"
                             f"Perhaps you meant to prepend an '@' character to a settings file, e.g. '@{inputs[0]}'?\n"
                             f"If not, note that only {ALLOWED_FILE_SUFFIXES_STRING} file types are supported")
        raise SystemExit(f"No valid files were found for processing. Note that currently only {ALLOWED_FILE_SUFFIXES_STRING} "
                         "file types are supported")
    output = opts["output"]
    if not output:
        outputfolders = [os.path.dirname(f) or "." for f in inputfiles]
    elif len(output) == len(inputfiles):
        outputfolders = list(output)
    elif len(output) == 1:
        outputfolders = [output[0]] * len(inputfiles)
    else:
        raise SystemExit(f"Incorrect number of output files passed ({len(output)}); must pass either 1 output folder (all "
                         "results are stored in this folder), or the same number of output folders as input image files "
                         f"({len(inputfiles)})")

    def paired(files, what):
        if not files:
            return [None] * len(inputfiles)
        if len(files) == len(inputfiles):
            return list(files)
        raise SystemExit(f"Number of {what} files passed ({len(files)}) does not equal the number of input image files "
                         f"passed ({len(inputfiles)})")

    maskfiles = paired(opts["mask"], "mask")
    if opts["B1map"]:
        assert opts["SetFlipAngle"] is None, "Cannot set a fixed flip angle using --SetFlipAngle when passing B1 maps using --B1map"
    b1files = paired(opts["B1map"], "B1 map")
    return [dict(inputfile=i, outputfolder=o, maskfile=m, B1mapfile=b, choppedinputfile=chop_allowed_suffix(os.path.basename(i)))
            for i, o, m, b in zip(inputfiles, outputfolders, maskfiles, b1files)]


def _kwargs(opts, fields, skip=()):
    kw = {}
    for k, v in opts.items():
        if v is None or (isinstance(v, list) and not v) or k not in fields or k in skip:
            continue
        kw[k] = tuple(v) if isinstance(v, list) else v
    return kw


def t2map_options(image, opts):
    kw = _kwargs(opts, T2MAP_FIELDS)
    for flag in ("SaveResidualNorm", "SaveDecayCurve", "SaveRegParam", "SaveNNLSBasis"):
        kw[flag] = bool(opts.get(flag))
    return T2mapOptions(MatrixSize=image.shape[:3], nTE=image.shape[3], ngpus=opts.get("ngpus", 0), Silent=opts.get("quiet", False), **kw)


def t2part_options(dist, opts):
    # nT2 must be explicitly passed, unless not performing T2-mapping, in which case it is inferred from `dist`
    kw = _kwargs(opts, T2PART_FIELDS, skip=() if opts["T2map"] else ("nT2",))
    kw.setdefault("nT2", dist.shape[3])
    return T2partOptions(MatrixSize=dist.shape[:3], Silent=opts.get("quiet", False), **kw)


def run_main(file_info, opts, log=print):
    """src/main.jl:333-420 for one input file.  Returns the dictionaries that were (or would be) written."""
    t_start = time.perf_counter()

    def timed(msg, f):
        t0 = time.perf_counter()
        r = f()
        log(f"{msg}: {time.perf_counter() - t0:.3f} seconds")
        return r

    image = timed(f"Loading input file: {file_info['inputfile']}", lambda: load_image(file_info["inputfile"], 4))
    if file_info["maskfile"] is not None:
        try:
            mask = timed(f"Applying mask from file: {file_info['maskfile']}", lambda: load_image(file_info["maskfile"], 3))
            image *= mask[..., None]
        except Exception as e:  # the reference warns and carries on (try_apply_maskfile!)
            warnings.warn(f"Error while loading mask file: {file_info['maskfile']}\n{e}")
    results = {}
    base = os.path.join(file_info["outputfolder"], file_info["choppedinputfile"])
    fused = None
    if opts["T2map"]:
        mopts = t2map_options(image, opts)
        b1 = None
        if file_info["B1mapfile"] is not None:
            try:
                b1 = timed(f"Loading B1 map from file: {file_info['B1mapfile']}", lambda: load_image(file_info["B1mapfile"], 3))
                assert b1.shape == image.shape[:3], "B1 map size must match the image matrix size"
            except Exception as e:
                warnings.warn(f"Error while loading B1 map file: {file_info['B1mapfile']}\n{e}")
                b1 = None
        popts = None
        if opts["T2part"]:
            popts = t2part_options(np.empty((*image.shape[:3], mopts.nT2)), opts)
        maps, dist = timed(f"Running T2mapSEcorr on file: {file_info['inputfile']}",
                           lambda: T2mapSEcorr(image, mopts, B1map=b1, t2part=popts))
        if popts is not None:
            fused = {k: maps.pop(k) for k in ("sfr", "sgm", "mfr", "mgm")}
        results["t2dist"], results["t2maps"] = {"dist": dist}, maps
        if not opts["dry"]:
            timed(f"Saving T2 distribution to file: {base}.t2dist.mat", lambda: save_mat(base + ".t2dist.mat", results["t2dist"]))
            timed(f"Saving T2 parameter maps to file: {base}.t2maps.mat", lambda: save_mat(base + ".t2maps.mat", maps))
    else:
        dist = image
    if opts["T2part"]:
        parts = fused if fused is not None else timed("Running T2partSEcorr", lambda: T2partSEcorr(dist, t2part_options(dist, opts)))
        results["t2parts"] = parts
        if not opts["dry"]:
            timed(f"Saving T2 parts maps to file: {base}.t2parts.mat", lambda: save_mat(base + ".t2parts.mat", parts))
    log(f"Finished ({time.perf_counter() - t_start:.2f} seconds)")
    return results


def main(args=None):
    """DECAES.main(args): process every input file; returns the list of result dictionaries."""
    opts = parse_cli(sys.argv[1:] if args is None else list(args))
    log = (lambda *a, **k: None) if opts["quiet"] else print
    out = []
    for info in get_file_infos(opts):
        if not opts["dry"]:
            os.makedirs(info["outputfolder"], exist_ok=True)
        out.append(run_main(info, opts, log))
    return out


if __name__ == "__main__":
    main()
