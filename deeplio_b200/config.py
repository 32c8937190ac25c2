"""Process-wide configuration container read by every net constructor.

Mirrors the reference's ``ConfigContainer`` / ``build_config_container`` / ``get_config_container``
(deeplio/models/misc.py:167-195): ``seq_size = len(combinations)``, ``timestamps = len(combinations[0])``,
``device`` and ``batch_size`` from the argparse namespace.  When this package is installed into the
reference (``deeplio_b200.install``), the reference's own container is used if it has been built, so
``train.py`` / ``test.py`` need no change.
"""
import sys

import numpy as np


class ConfigContainer:
    def __init__(self, cfg, args):
        self.cfg = cfg
        self.args = args
        self.ds_cfg = cfg["datasets"]
        self.curr_dataset_cfg = cfg["datasets"][cfg["current-dataset"]]
        self.combinations = np.array(self.ds_cfg["combinations"])
        self.seq_size = len(self.combinations)
        self.timestamps = len(self.combinations[0])
        self.device = args.device
        self.batch_size = args.batch_size
        self.seq_size_data = self.ds_cfg["sequence-size"]


_container = None


def build_config_container(cfg, args):
    global _container
    _container = ConfigContainer(cfg, args)
    return _container


def get_config_container():
    ref = sys.modules.get("deeplio.models.misc")
    if ref is not None and getattr(ref, "config_container", None) is not None:
        return ref.config_container
    if _container is None:
        raise ValueError("Config container must be created by Worker first!")
    return _container
