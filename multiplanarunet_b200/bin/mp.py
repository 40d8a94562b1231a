"""`mp <script> [args]` dispatcher (mirror of mpunet/bin/mp.py:13-55): scripts are the modules of
this package exposing entry_func(args)."""
import argparse
import importlib
import os
import sys


def get_parser():
    here = os.path.dirname(os.path.abspath(__file__))
    mods = sorted(m[:-3] for m in os.listdir(here) if m.endswith(".py") and m not in ("mp.py", "__init__.py"))
    usage = ("mp [script] [script args...]\n\nMulti-Planar UNet on B200\n--------------------------\n"
             "Available scripts:\n- " + "\n- ".join(mods))
    parser = argparse.ArgumentParser(usage=usage)
    parser.add_argument("script", help="Name of the mp script to run.", choices=mods)
    parser.add_argument("args", help="Arguments passed to script", nargs=argparse.REMAINDER)
    return parser


def entry_func(args=None):
    parsed = get_parser().parse_args(sys.argv[1:] if args is None else args)
    mod = importlib.import_module("multiplanarunet_b200.bin." + parsed.script)
    mod.entry_func(parsed.args)


if __name__ == "__main__":
    entry_func()
