"""Shared helpers of the parity tests: run the same Scene through the CUDA product and the CPU oracle, compare."""
import numpy as np

from openclrenderer_b200 import Renderer
from oracle.binding import Oracle


def render_both(scene, frames=1, shadows=True, threads=0, cfg=None):
    cfg = cfg or scene.cfg
    g = Renderer(cfg)
    o = Oracle(cfg, threads=threads)
    scene.upload(g)
    scene.upload(o)
    scene.render(g, frames=frames, shadows=shadows)
    scene.render(o, frames=frames, shadows=shadows)
    return g, o


def colour_stats(a, b, rows=None):
    """a, b: (H,W,4) uint8. Returns (fraction within +-1 LSB on all channels, max abs channel difference)."""
    if rows is not None:
        a, b = a[rows[0]:rows[1]], b[rows[0]:rows[1]]
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    within1 = (d.max(axis=-1) <= 1).mean()
    return float(within1), int(d.max())


def assert_frame_parity(g, o, rows=None, min_within1=0.999, max_lsb=2, check_records=True, label=""):
    """north_star bar: depth and id buffers bit-exact; RGBA8 within +-1 LSB on >= 99.9 % of pixels and +-2 LSB max."""
    gd, od = g.read_depth(), o.read_depth()
    gi, oi = g.read_ids(), o.read_ids()
    sl = slice(None) if rows is None else slice(rows[0], rows[1])
    nd = int((gd[sl] != od[sl]).sum())
    assert nd == 0, f"{label}: {nd} depth pixels differ"
    covered = od[sl] != 0xFFFFFFFF
    ni = int((gi[sl][covered] != oi[sl][covered]).sum())
    assert ni == 0, f"{label}: {ni} id pixels differ"
    if check_records:
        gf, of = g.read_fragments(), o.read_fragments()
        assert gf.shape == of.shape, f"{label}: fragment count {gf.shape} vs {of.shape}"
        assert np.array_equal(gf, of), f"{label}: fragment records differ"
        gc, oc = g.read_cutdown(), o.read_cutdown()
        assert gc.shape == oc.shape, f"{label}: cutdown count {gc.shape} vs {oc.shape}"
        # slots of culled triangles are never written (cl2.cl:4361) -> compare only the slots fragments refer to
        used = np.unique(of[:, 2]) if len(of) else np.zeros(0, np.int64)
        assert np.array_equal(gc[used].view(np.uint32), oc[used].view(np.uint32)), f"{label}: projected triangles differ"
    within1, mx = colour_stats(g.read_rgba8(), o.read_rgba8(), rows)
    assert within1 >= min_within1, f"{label}: only {within1 * 100:.4f}% of pixels within +-1 LSB"
    assert mx <= max_lsb, f"{label}: max channel difference {mx} LSB"
    return {"within1": within1, "max_lsb": mx, "covered": int(covered.sum())}
