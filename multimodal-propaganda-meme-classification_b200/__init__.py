"""b200mm: B200-native (sm_100a) train/infer engine for the task-2C late-fusion propaganda-meme classifier.

Import as ``b200mm`` (a shim at the repo root maps that name onto this directory, whose name follows the
reference repository and is not a valid Python identifier).  Public surface = the reference's own:

    MultimodalClassifier(num_classes)            example_scripts/Multimodal_example_task2C.txt:152-197
    train / test / evaluate                       ...txt:200-242, 259-280
    CrossEntropyLoss, FusedAdam                   ...txt:248-249
    ensemble.*                                    example_scripts/combine_preds.py
    setup(k) / run_folds / combine_folds          example_scripts/Multimodal_example_task2C.py:50-192, 882-885
    ConvNeXtTiny, BertPoolerModel, extract_features   baselines/extract_feat.py:52-67, 103-112
    decode_jpeg / collate_jpeg                    the Dataset's Image.open(path).convert("RGB") (.txt:50; .py:270), split host / GPU
"""
__version__ = "0.1.0"

_LAZY = {
    "MultimodalClassifier": ("model", "MultimodalClassifier"),
    "MultimodalClassifierHEAD": ("head_model", "MultimodalClassifierHEAD"),
    "TextConfig": ("text_tower", "TextConfig"),
    "ImageConfig": ("image_tower", "ImageConfig"),
    "ViTConfig": ("vit_tower", "ViTConfig"),
    "FusedAdam": ("optim", "FusedAdam"),
    "get_linear_schedule_with_warmup": ("optim", "get_linear_schedule_with_warmup"),
    "CrossEntropyLoss": ("loop", "CrossEntropyLoss"),
    "SigmoidFocalLoss": ("loop", "SigmoidFocalLoss"),
    "train": ("loop", "train"),
    "test": ("loop", "test"),
    "evaluate": ("loop", "evaluate"),
    "predict": ("loop", "predict"),
    "get_features": ("loop", "get_features"),
    "seed_everything": ("loop_head", "seed_everything"),
    "stratified_kfold": ("loop_head", "stratified_kfold"),
    "get_params": ("loop_head", "get_params"),
    "GraphedTrainStep": ("graph", "GraphedTrainStep"),
    "ConvNeXtTiny": ("features", "ConvNeXtTiny"),
    "BertPoolerModel": ("features", "BertPoolerModel"),
    "extract_features": ("features", "get_features"),
    "write_features_json": ("features", "write_features_json"),
    "setup": ("folds", "setup"),
    "run_folds": ("folds", "run_folds"),
    "combine_folds": ("folds", "combine_folds"),
    "decode_jpeg": ("jpeg", "decode_jpeg"),
    "collate_jpeg": ("jpeg", "collate_jpeg"),
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        return getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
    raise AttributeError(name)
