"""Importable alias of the package directory `2048_q-learning_b200/` (not a valid Python identifier)."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("2048_q-learning_b200")
