"""multiplanarunet_b200.utils package."""
