"""Host-side ROTModifyParentsOnce helpers (graph-state/src/rooted_tree/space.rs) used only to seed a host model."""
import numpy as np

from . import capi


def write_vec(n, parents, mask):  # rooted_tree/space.rs:91-101
    a = capi.action_dim(n)
    v = np.zeros(2 * a, dtype=np.float32)
    for c in range(2, n - 1):
        v[c * (c - 1) // 2 + int(parents[c]) - 1] = 1.0
    for w, word in enumerate(np.asarray(mask, dtype=np.uint32)):
        word = int(word)
        while word:
            b = (word & -word).bit_length() - 1
            v[a + w * 32 + b] = 1.0
            word &= word - 1
    return v
