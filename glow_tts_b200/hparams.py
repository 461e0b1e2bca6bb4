"""Hyper-parameters.  The reference parses ./Hyper_Parameters.yaml from the CWD at
import time into a nested argparse.Namespace called ``hp`` (Arg_Parser.py:3-12,
Modules.py:9-13) and every module constructor reads that global.  The same keys
are honoured here: ``load_hparams()`` reads the same YAML (if present) over the
defaults below, which are the reference's shipped values for the keys the hot
path reads (Hyper_Parameters.yaml:3-57,112-121)."""
import argparse
import copy
import os

import yaml

DEFAULTS = {
    "Sound": {"Mel_Dim": 80, "Max_Abs_Mel": 4},
    "Use_Cython_Alignment": True,
    "Mode": "Vanilla",
    "Encoder": {
        "Channels": 192, "Embedding_Tokens": 35,
        "Prenet": {"Kernel_Size": 5, "Dropout_Rate": 0.5, "Stacks": 3},
        "Transformer": {
            "Attention": {"Heads": 2, "Window_Size": 4},
            "Conv": {"Kernel_Size": 3, "Calc_Channels": 768},
            "Dropout_Rate": 0.1, "Stacks": 6},
        "Duration_Predictor": {"Kernel_Size": 3, "Channels": 256, "Stacks": 2, "Dropout_Rate": 0.1},
    },
    "Decoder": {
        "Stack": 12, "Num_Squeeze": 2, "Num_Split": 4,
        "Affine_Coupling": {"Calc_Channels": 192,
                            "WaveNet": {"Num_Layers": 4, "Kernel_Size": 5, "Dropout_Rate": 0.05}},
    },
    "Speaker_Embedding": {"Type": "LUT", "Num_Speakers": 109, "Embedding_Size": 256},
    "Train": {
        "Batch_Size": 32,
        "Train_Pattern": {"Mel_Length": {"Min": 50, "Max": 1000}, "Text_Length": {"Min": 10, "Max": 200}},
        "Learning_Rate": {"Initial": 1.0e-3, "Base": 4000},
        "ADAM": {"Beta1": 0.9, "Beta2": 0.999, "Epsilon": 1.0e-6},
        "Weight_Decay": 1.0e-6, "Gradient_Norm": 5.0,
    },
    "Device": "0",
    # not in the reference: which arithmetic the sm_100a core uses ('bf16' | 'fp32')
    "Precision": "bf16",
}


def _merge(base, over):
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(base.get(k), dict):
            _merge(base[k], v)
        else:
            base[k] = v
    return base


def to_namespace(d):
    ns = argparse.Namespace()
    for k, v in d.items():
        setattr(ns, k, to_namespace(v) if isinstance(v, dict) else v)
    return ns


def load_hparams(path=None, **overrides):
    """Namespace with the reference's key layout.  path=None -> ./Hyper_Parameters.yaml
    if it exists, else the defaults.  Keyword overrides use the top-level keys
    (e.g. Mode='SE') or dotted paths (**{'Decoder.Stack': 2})."""
    d = copy.deepcopy(DEFAULTS)
    if path is None and os.path.exists("Hyper_Parameters.yaml"):
        path = "Hyper_Parameters.yaml"
    if path is not None:
        with open(path, encoding="utf-8") as f:
            _merge(d, yaml.load(f, Loader=yaml.Loader) or {})
    for key, val in overrides.items():
        node = d
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = val
    return to_namespace(d)
