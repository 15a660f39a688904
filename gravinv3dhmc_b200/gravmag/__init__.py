"""Sensitivity-kernel builders (prism / tesseroid gz) and the wavelet compressors."""
