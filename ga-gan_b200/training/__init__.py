"""Callers of the hot path: networks (incl. `modulated_conv2d`), loss phases, the training step."""
