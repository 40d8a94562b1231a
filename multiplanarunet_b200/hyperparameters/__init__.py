from .hparams import YAMLHParams  # noqa: F401
