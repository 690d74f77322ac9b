"""On-disk formats either side of the prediction path (host only)."""
