"""B200-native drop-ins for the reference's model/ directory (unet.py, inference.py, evaluation*.py)."""
