"""Built-in skeleton graph definitions.

Same graphs as the reference's ``data/skeletons/*.json`` (joint count, bone
list, centre joint), expressed as compact bone strings; ``skeleton(name)``
expands one into the ``{num_node, edge, center}`` dict the reference passes to
``Graph(**graph)`` (processor.py:166): every self-loop first, then the bones.
``load_graph`` also accepts a path to a reference-format JSON file.
"""
import json
import os

# name -> (num_node, center, "a-b a-b ...")
_BONES = {
    'pku-mmd': (25, 20,
                "0-1 1-20 2-20 3-2 4-20 5-4 6-5 7-6 8-20 9-8 10-9 11-10 12-0 "
                "13-12 14-13 15-14 16-0 17-16 18-17 19-18 21-7 22-7 23-11 24-11"),
    'ntu-rgb+d': (25, 20,
                  "0-1 1-20 2-20 3-2 4-20 5-4 6-5 7-6 8-20 9-8 10-9 11-10 12-0 "
                  "13-12 14-13 15-14 16-0 17-16 18-17 19-18 21-7 22-7 23-11 24-11"),
    'ntu-edge': (24, 2,
                 "0-1 2-1 3-2 4-1 5-4 6-5 7-6 8-1 9-8 10-9 11-10 12-0 13-12 "
                 "14-13 15-14 16-0 17-16 18-17 19-18 20-21 21-7 22-23 23-11"),
    'imu_fogit_ABCD': (7, 0, "0-1 1-2 2-3 0-4 4-5 5-6"),
    'hugadb': (6, 0, "1-0 2-1 3-0 4-3 5-0"),
    'tp-vicon': (9, 0, "1-0 2-1 3-2 4-3 5-0 6-5 7-6 8-7"),
    'lara': (19, 0,
             "1-0 2-1 3-2 4-3 5-0 6-5 7-6 8-7 9-0 10-9 11-9 12-10 13-12 14-13 "
             "15-9 16-15 17-16 18-17"),
    'openpose': (18, 1,
                 "4-3 3-2 7-6 6-5 13-12 12-11 10-9 9-8 11-5 8-2 5-1 2-1 0-1 "
                 "15-0 14-0 17-15 16-14"),
    'coco': (17, 0,
             "15-13 13-11 16-14 14-12 11-12 5-11 6-12 5-6 7-5 8-6 9-7 10-8 "
             "1-2 1-0 2-0 3-1 4-2 3-5 4-6"),
}


def names():
    return sorted(_BONES)


def skeleton(name):
    """``{num_node, edge, center}`` for a built-in graph."""
    if name not in _BONES:
        raise KeyError("unknown skeleton %r (have: %s)" % (name, ', '.join(names())))
    v, center, bones = _BONES[name]
    edge = [[i, i] for i in range(v)]
    edge += [[int(a), int(b)] for a, b in (p.split('-') for p in bones.split())]
    return {'num_node': v, 'edge': edge, 'center': center}


def load_graph(name_or_path):
    """Built-in name, or path to a ``{num_node, edge, center}`` JSON file."""
    if isinstance(name_or_path, dict):
        return name_or_path
    if os.path.isfile(name_or_path):
        with open(name_or_path) as f:
            return json.load(f)
    return skeleton(name_or_path)
