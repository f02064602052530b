"""Configuration loading shared by the drivers.

The reference reads ``../tools/configs.yml`` and ``./user_configs.yml`` relative to the
current directory (pytorch_quantizer.py:23-32, reconstruction.py:104-105).  That still works;
additionally a dict or a path can be passed explicitly, and the packaged configs.yml is the
default when no cwd-relative file exists."""
import copy
import os

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))

DEFAULT_USER_CONFIG = {
    "PATH": {"DATA_PATH": "", "MODEL_NET_PATH": "", "MODEL_PATH": "",
             "QUANTITY_MODEL_PATH": "./workdir/quantity_model.pth"},
    "MODEL": {"INPUT_SHAPE": "1,3,224,224"},
    "PRE_PROCESS": {"IMG": 1, "IMG_SET": {"MEAN": 128, "RESIZE": "224,224", "SCALE": 0.0075}},
    "SETTINGS": {"DEVICE": "gpu", "GPU": 0},
}


def _read(path):
    with open(path) as f:
        return yaml.safe_load(f)


def load_tool_config(config=None):
    if isinstance(config, dict):
        return copy.deepcopy(config)
    if isinstance(config, str):
        return _read(config)
    if os.path.isfile("../tools/configs.yml"):
        return _read("../tools/configs.yml")
    return _read(os.path.join(_HERE, "configs.yml"))


def load_user_config(user_config=None):
    if isinstance(user_config, dict):
        return copy.deepcopy(user_config)
    if isinstance(user_config, str):
        return _read(user_config)
    if os.path.isfile("./user_configs.yml"):
        return _read("./user_configs.yml")
    return copy.deepcopy(DEFAULT_USER_CONFIG)
