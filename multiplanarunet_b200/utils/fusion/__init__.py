from .fuse_and_predict import (predict_volume, map_real_space_pred, predict_and_map,
                               predict_multi_view, voxel_grid_center, stack_collections)  # noqa: F401
