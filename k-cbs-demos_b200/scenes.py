"""Scene constants of the reference demos (src/demo_*.cpp), loaded from scenes.json, which
tools/extract_scenes.py generated from the reference tree.  Robot index order follows
addIndividual order: std::map key order where the demo uses std::map (Corridor, Congested), source
listing order where it uses std::unordered_map (Empty scenes; the true order is implementation
defined)."""
from __future__ import annotations

import json
import os
from typing import Dict, List

import numpy as np

from .capi import World

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scenes.json")
with open(_PATH) as _f:
    SCENES: Dict[str, dict] = json.load(_f)


def names() -> List[str]:
    return sorted(SCENES)


def world(name: str) -> World:
    s = SCENES[name]
    dims = [[s["robot_length"], s["robot_width"]] for _ in s["robots"]]
    return World(x_max=s["x_max"], y_max=s["y_max"], robot_dims=dims, obstacles=s["obstacles"])


def start_states(name: str) -> np.ndarray:
    """[5, R] float32 start states (sx, sy, 0, 0, 0), demo_*.cpp:121-126."""
    s = SCENES[name]
    out = np.zeros((5, len(s["robots"])), np.float32)
    for i, r in enumerate(s["robots"]):
        out[0, i], out[1, i] = r["start"]
    return out


def goals(name: str) -> np.ndarray:
    s = SCENES[name]
    return np.asarray([r["goal"] for r in s["robots"]], np.float32)
