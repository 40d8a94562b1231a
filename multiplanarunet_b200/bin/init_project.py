"""`mp init_project` (mirror of mpunet/bin/init_project.py:5-86): copy the preset YAML(s) into a new
project folder and point the {train,val,test,aug}_data.base_dir entries at <data_dir>/<split>."""
import os
from argparse import ArgumentParser
from glob import glob

from ..hyperparameters import YAMLHParams

DEFAULTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "defaults")


def get_parser():
    parser = ArgumentParser(description="Create a new project folder")
    parser.add_argument("--name", type=str, required=True, help="the name of the project folder")
    parser.add_argument("--root", type=str, default=os.path.abspath("./"),
                        help="a path to the root folder in which the project will be initialized")
    parser.add_argument("--model", type=str, default="MultiPlanar",
                        help="model type parameter file (%s)" % ", ".join(sorted(os.listdir(DEFAULTS))))
    parser.add_argument("--data_dir", type=str, default=None, help="Root data folder for the project")
    return parser


def copy_yaml_and_set_data_dirs(in_path, out_path, data_dir):
    hp = YAMLHParams(in_path, no_log=True, no_version_control=True)
    for split in ("train", "val", "test", "aug"):
        d = (os.path.abspath(data_dir) + "/" + split) if data_dir else "Null"
        hp.set_value(split + "_data", "base_dir", d, overwrite=True)
    hp.save_current(out_path)


def entry_func(args=None):
    a = get_parser().parse_args(args)
    folder = os.path.join(os.path.abspath(a.root), a.name)
    preset = os.path.join(DEFAULTS, a.model)
    if not os.path.isdir(preset):
        raise ValueError("Unknown model preset '%s'" % a.model)
    if not os.path.exists(folder):
        os.makedirs(folder)
    elif os.listdir(folder):
        raise OSError("Folder at '%s' already exists and is not empty." % folder)
    for p in glob(os.path.join(preset, "*.yaml")):
        copy_yaml_and_set_data_dirs(p, os.path.join(folder, os.path.basename(p)), a.data_dir)
    print("Project initialised at %s" % folder)
