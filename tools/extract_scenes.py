#!/usr/bin/env python
"""Extracts the scenario constants (start/goal maps, obstacles, robot size, workspace bounds, merge
bound, time-outs) of every demo main in /root/reference/src into k-cbs-demos_b200/scenes.json.
Data only -- no reference code is copied.  Run in the build container (the reference tree is not
present on the GPU box); the JSON it writes is committed."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "k-cbs-demos_b200", "scenes.json")

num = r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?"


def parse_map(text, name):
    m = re.search(name + r"\s*\{(.*?)\};", text, re.S)
    if not m:
        return None, None
    ordered = "unordered_map" not in text[max(0, m.start() - 90):m.start()]
    items = re.findall(r'\{\s*"([^"]+)"\s*,\s*\{\s*(%s)\s*,\s*(%s)\s*\}\s*\}' % (num, num), m.group(1))
    return [(k, float(a), float(b)) for k, a, b in items], ordered


scenes = {}
for fn in sorted(os.listdir(REF)):
    if not fn.endswith(".cpp"):
        continue
    text = open(os.path.join(REF, fn)).read()
    # drop commented-out lines so disabled obstacles / robots are not picked up
    live = "\n".join(l for l in text.splitlines() if not l.strip().startswith("//"))
    starts, ordered = parse_map(live, "start_map")
    goals, _ = parse_map(live, "goal_map")
    obs_re = r"RectangularObstacle\(\s*(%s)\s*,\s*(%s)\s*,\s*(%s)\s*,\s*(%s)\s*\)" % (num, num, num, num)
    obstacles = [[float(v) for v in m] for m in re.findall(obs_re, live)]
    rob = re.search(r"RectangularRobot\([^,]+,\s*(" + num + r")\s*,\s*(" + num + r")\s*\)", live)
    space = re.search(r"createBounded2ndOrderCarStateSpace\(\s*(\d+)\s*,\s*(\d+)\s*\)", live)
    merge = re.search(r"set(?:KCBS)?MergeBound\(\s*(\d+)\s*\)", live)
    solve = re.search(r"solve\(\s*(" + num + r")\s*\)", live) or re.search(r"setSolveTime\(\s*(" + num + r")\s*\)", live)
    low = re.search(r"setLowLevelSolveTime\(\s*(" + num + r")\s*\)", live)
    name = fn[len("demo_"):-len("_dyn2ndOrderCars.cpp")]
    goal_of = {k: (a, b) for k, a, b in goals}
    scenes[name] = {
        "source": "src/" + fn,
        "x_max": int(space.group(1)), "y_max": int(space.group(2)),
        "robot_length": float(rob.group(1)), "robot_width": float(rob.group(2)),
        # std::map iterates in key order; std::unordered_map order is implementation defined, so
        # for those scenes the listed (source) order is used as the robot index order.
        "ordered_map": bool(ordered),
        "robots": [{"name": k, "start": [a, b], "goal": list(goal_of[k])} for k, a, b in
                   (sorted(starts) if ordered else starts)],
        "obstacles": obstacles,
        "merge_bound": int(merge.group(1)) if merge else None,
        "solve_time": float(solve.group(1)) if solve else None,
        "low_level_solve_time": float(low.group(1)) if low else 5.0,
    }

with open(OUT, "w") as f:
    json.dump(scenes, f, indent=1, sort_keys=True)
for k, v in scenes.items():
    print(f"{k:45s} {v['x_max']}x{v['y_max']} robots={len(v['robots'])} obstacles={len(v['obstacles'])} "
          f"robot={v['robot_length']}x{v['robot_width']} merge={v['merge_bound']} solve={v['solve_time']}")
