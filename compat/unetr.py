"""Drop-in for the reference's local `unetr.py`: put this directory (and the repo root) on PYTHONPATH and
`from unetr import UNETR` (unetr_ranking_pretraining_3d.py:34) resolves to the B200 implementation."""
import importlib

UNETR = importlib.import_module("3dmedicalimagesegmentation_b200").UNETR
