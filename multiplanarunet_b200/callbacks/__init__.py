from .callbacks import (Callback, CSVLogger, DelayedCallback, DividerLine, EarlyStopping,  # noqa: F401
                        FGBatchBalancer, ModelCheckPointClean, ReduceLROnPlateau, TensorBoard, TrainTimer)
from .funcs import init_callback_objects, remove_validation_callbacks  # noqa: F401
from .validation import Validation  # noqa: F401
