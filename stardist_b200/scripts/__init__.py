"""Command line entry points (stardist/scripts/predict2d.py, predict3d.py)."""
