"""Shared helpers of the parity tests: the same scene + UBO through the CUDA C ABI and through the CPU oracle."""
import numpy as np


def scene_for(oit, **kw):
    st = oit.State(**kw)
    verts, idx, ipo = oit.generate_scene(st)
    return st, verts, idx, ipo


def oracle_cfg(O, st, W, H):
    return O.make_config(algorithm=st.algorithm, oitLayers=st.oitLayers,
                         linkedListAllocatedPerElement=st.linkedListAllocatedPerElement,
                         percentTransparent=st.percentTransparent, tailBlend=int(st.tailBlend),
                         interlockIsOrdered=int(st.interlockIsOrdered), numObjects=st.numObjects, subdiv=st.subdiv,
                         scaleMin=st.scaleMin, scaleWidth=st.scaleWidth, aaType=st.aaType, width=W, height=H)


def make_oracle(O, st, W, H, verts, idx, ipo, ubo, threads=1):
    o = O.Oracle(oracle_cfg(O, st, W, H), threads=threads)
    o.set_scene(verts, idx, ipo)
    sd = O.SceneData.from_buffer_copy(bytes(ubo))
    return o, sd


def walk_lists(abuf, heads):
    """Linked-list A-buffer -> per-pixel tuple of (colour, depth, mask) in list order (node numbering removed)."""
    nodes = abuf.reshape(-1, 4)
    out = {}
    for p in np.nonzero(heads)[0]:
        off, lst = int(heads[p]), []
        while off:
            c, d, m, nxt = nodes[off]
            lst.append((int(c), int(d), int(m)))
            off = int(nxt)
        out[int(p)] = tuple(lst)
    return out


def max_channel_diff(a, b):
    return int(np.abs(a.view(np.uint8).astype(np.int32) - b.view(np.uint8).astype(np.int32)).max())
