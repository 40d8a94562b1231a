"""multiplanarunet_b200.bin package."""
