"""Launcher: run an unchanged reference script on the B200 path.

    torchrun --nproc-per-node 8 -m diffusion_by_maxentirl_b200.run generate_large.py --log_dir ... --batchsize 64

The reference checkout must be the working directory (or on PYTHONPATH); `install()` rebinds its hot-path classes to
the drop-ins before the script's own imports run."""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m diffusion_by_maxentirl_b200.run <reference script.py> [script args...]")
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)) or os.getcwd())
    import diffusion_by_maxentirl_b200 as pkg

    bound = pkg.install()
    print(f"[dxmi_b200] bound {len(bound)} hot-path symbols to libdxmi_b200.so", file=sys.stderr)
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
