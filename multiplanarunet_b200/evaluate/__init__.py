from .metrics import compute_dice, dice, dice_all, label_counts  # noqa: F401
