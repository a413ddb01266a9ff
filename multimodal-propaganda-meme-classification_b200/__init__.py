"""b200mm: B200-native (sm_100a) train/infer engine for the task-2C late-fusion propaganda-meme classifier.

Import as ``b200mm`` (a shim at the repo root maps that name onto this directory, whose name follows the
reference repository and is not a valid Python identifier).
"""
__version__ = "0.1.0"
