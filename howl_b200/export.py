"""Checkpoint export for the browser runtime (Honkling).

Mirrors ``training/run/export_honkling.py:19-30`` of the reference: the state dict is written as a JavaScript assignment
``weights['<NAME>'] = {key: nested lists, ...}`` in state-dict order; for ``RES8`` three constant ``scaleN.scale`` vectors of
45 ones are appended (Honkling's res8 has learnable per-channel scales that howl's does not).  The text is byte-identical to
the reference script's output for the same state dict (``tests/test_host_logic.py``).
"""
import json
from collections import OrderedDict
from typing import Mapping

import torch


def honkling_dict(state_dict: Mapping[str, torch.Tensor], name: str) -> "OrderedDict[str, list]":
    sd = OrderedDict(state_dict)
    if name == "RES8":
        for k in ("scale1.scale", "scale3.scale", "scale5.scale"):
            sd[k] = torch.ones(45)
    return OrderedDict((k, torch.as_tensor(v).detach().cpu().tolist()) for k, v in sd.items())


def export_honkling(state_dict: Mapping[str, torch.Tensor], name: str) -> str:
    """The contents of the ``.js`` file the reference's exporter writes."""
    return f"weights['{name}'] = " + json.dumps(honkling_dict(state_dict, name))


def write_honkling(state_dict: Mapping[str, torch.Tensor], name: str, path: str) -> None:
    with open(path, "w") as f:
        f.write(export_honkling(state_dict, name))
