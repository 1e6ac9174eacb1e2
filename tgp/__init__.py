"""Namespace package root; the product lives in `tgp.pytorch_b200`."""
