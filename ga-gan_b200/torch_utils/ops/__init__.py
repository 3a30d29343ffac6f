"""Operator API of the StyleGAN2 conv hot path, served by libgagan_b200.so (sm_100a)."""
