"""oracle package."""
