"""Minimal stand-in for ``posepile.joint_info.JointInfo`` (un-vendored dependency of the reference,
multiperson_model.py:4,25): names, edges, ``n_joints`` and the left/right ``mirror_mapping`` used by the TTA un-flip
(multiperson_model.py:249).  posepile's convention: a joint whose name starts with 'l' mirrors to the same name with 'r'."""
import numpy as np


class JointInfo:
    def __init__(self, names, edges, mirror_mapping=None):
        self.names = [str(n) for n in names]
        self.stick_figure_edges = [tuple(int(i) for i in e) for e in np.asarray(edges).reshape(-1, 2)]
        self.n_joints = len(self.names)
        if mirror_mapping is None:
            idx = {n: i for i, n in enumerate(self.names)}
            mirror_mapping = []
            for i, n in enumerate(self.names):
                other = ('r' if n[0] == 'l' else 'l' if n[0] == 'r' else n[0]) + n[1:] if n else n
                mirror_mapping.append(idx.get(other, i) if n and n[0] in 'lr' else i)
        self.mirror_mapping = [int(i) for i in mirror_mapping]
