from .trainer import fit_loop  # noqa: F401
