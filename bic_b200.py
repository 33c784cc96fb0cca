"""Importable alias for the hyphenated package directory `binary-image-compression_b200/`."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("binary-image-compression_b200")
