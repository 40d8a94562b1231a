"""`mp predict_3D` - CLI name kept for drop-in compatibility (mpunet/bin/predict_3D.py:12-39).

In the reference this script serves the 3D-UNet patch/box model family (predict_3D.py:176-193), which
is outside the MultiPlanar hot path this implementation covers; the 6-view predict + fusion path is
`mp predict` (bin/predict.py:294-366), exactly as in the reference."""
from argparse import ArgumentParser


def get_argparser():
    p = ArgumentParser(description="Predict using a mpunet 3D model (not part of the B200 hot path).")
    # the reference's flag set (mpunet/bin/predict_3D.py:12-39), so that existing command lines parse
    p.add_argument("--project_dir", type=str, default="./")
    p.add_argument("-f", help="Predict on a single file")
    p.add_argument("-l", help="Optional single label file to use with -f")
    p.add_argument("--data_dir", type=str, default=None)
    p.add_argument("--out_dir", type=str, default="predictions")
    p.add_argument("--num_GPUs", type=int, default=1)
    p.add_argument("--overwrite", action="store_true")
    p.add_argument("--no_eval", action="store_true")
    p.add_argument("--strides", type=int, default=None)
    p.add_argument("--extra", default="2x")
    p.add_argument("--force_GPU", type=int, default=-1)
    p.add_argument("--save_only_pred", action="store_true")
    return p


def entry_func(args=None):
    get_argparser().parse_args(args)
    raise NotImplementedError(
        "mp predict_3D drives the 3D-UNet model family, which is not implemented on the B200 path; "
        "MultiPlanar projects (model_class_name: UNet) predict with `mp predict`.")
