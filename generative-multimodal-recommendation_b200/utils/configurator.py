"""YAML configuration in the reference's layout and merge order.

Mirrors ``Config`` of GenMMRec/src/utils/configurator.py:46-129: ``overall.yaml`` -> ``dataset/<d>.yaml`` ->
``model/<M>.yaml`` (-> ``mg.yaml``) -> the ``config_dict`` argument, ``hyper_parameters`` collected
across files, missing keys read as ``None``.  The directory is this package's ``configs/`` by default;
point ``config_dir`` (or ``GMR_CONFIG_DIR``) at the reference's ``src/configs`` to consume its files
unchanged.
"""
import os
import re

import torch
import yaml

_PKG_CONFIGS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs")


def _yaml_loader():
    # scientific notation without a dot ("1e-05") must parse as float, as in the reference's loader
    loader = yaml.FullLoader
    loader.add_implicit_resolver(
        u"tag:yaml.org,2002:float",
        re.compile(u"""^(?:
         [-+]?(?:[0-9][0-9_]*)\\.[0-9_]*(?:[eE][-+]?[0-9]+)?
        |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
        |\\.[0-9_]+(?:[eE][-+][0-9]+)?
        |[-+]?[0-9][0-9_]*(?::[0-5]?[0-9])+\\.[0-9_]*
        |[-+]?\\.(?:inf|Inf|INF)
        |\\.(?:nan|NaN|NAN))$""", re.X),
        list(u"-+0123456789."))
    return loader


class Config(object):
    def __init__(self, model=None, dataset=None, config_dict=None, mg=False, config_dir=None):
        config_dict = dict(config_dict or {})
        config_dict["model"] = model
        config_dict["dataset"] = dataset
        self.config_dir = config_dir or os.environ.get("GMR_CONFIG_DIR") or _PKG_CONFIGS
        self.final_config_dict = self._load_files(config_dict, mg)
        self.final_config_dict.update(config_dict)
        self._set_default_parameters()
        self._init_device()

    def _load_files(self, config_dict, mg):
        files = [os.path.join(self.config_dir, "overall.yaml"),
                 os.path.join(self.config_dir, "dataset", "{}.yaml".format(config_dict["dataset"])),
                 os.path.join(self.config_dir, "model", "{}.yaml".format(config_dict["model"]))]
        if mg:
            files.append(os.path.join(self.config_dir, "mg.yaml"))
        merged, hyper = {}, []
        for f in files:
            if os.path.isfile(f):
                with open(f, "r", encoding="utf-8") as fh:
                    data = yaml.load(fh.read(), Loader=_yaml_loader()) or {}
                if data.get("hyper_parameters"):
                    hyper.extend(data["hyper_parameters"])
                merged.update(data)
        merged["hyper_parameters"] = hyper
        return merged

    def _set_default_parameters(self):
        smaller = ["rmse", "mae", "logloss"]
        vm = (self.final_config_dict.get("valid_metric") or "Recall@20").split("@")[0]
        self.final_config_dict["valid_metric_bigger"] = vm not in smaller
        if "seed" not in self.final_config_dict["hyper_parameters"]:
            self.final_config_dict["hyper_parameters"] += ["seed"]

    def _init_device(self):
        if "device" in self.final_config_dict and self.final_config_dict["device"] is not None:
            self.final_config_dict["device"] = torch.device(self.final_config_dict["device"])
            return
        use_gpu = self.final_config_dict.get("use_gpu", True)
        # one process per GPU: honour LOCAL_RANK instead of rewriting CUDA_VISIBLE_DEVICES
        if torch.cuda.is_available() and use_gpu:
            self.final_config_dict["device"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
        else:
            self.final_config_dict["device"] = torch.device("cpu")

    def __setitem__(self, key, value):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        self.final_config_dict[key] = value

    def __getitem__(self, item):
        return self.final_config_dict.get(item, None)

    def __contains__(self, key):
        if not isinstance(key, str):
            raise TypeError("index must be a str.")
        return key in self.final_config_dict

    def __str__(self):
        return "\n" + "\n".join("{}={}".format(k, v) for k, v in self.final_config_dict.items()) + "\n\n"

    __repr__ = __str__
