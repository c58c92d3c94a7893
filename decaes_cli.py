#!/usr/bin/env python
"""`decaes` command line (mirror of DECAES.main, src/main.jl) over libdecaes_cuda:

    python decaes_cli.py image.nii.gz --T2map --T2part --TE 7e-3 --nT2 40 --T2Range 10e-3 2.0 \
        --SPWin 10e-3 25e-3 --MPWin 25e-3 200e-3 --Reg lcurve --output results/
    python decaes_cli.py @settings.txt

The package directory is called `decaes.jl_b200` (a dot in its name), so it is loaded by path here."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))


def load_package():
    name = "decaes_jl_b200"
    if name in sys.modules:
        return sys.modules[name]
    pkg_dir = os.path.join(ROOT, "decaes.jl_b200")
    spec = importlib.util.spec_from_file_location(name, os.path.join(pkg_dir, "__init__.py"), submodule_search_locations=[pkg_dir])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main(args=None):
    load_package()
    import importlib
    return importlib.import_module("decaes_jl_b200.cli").main(args)


if __name__ == "__main__":
    main()
