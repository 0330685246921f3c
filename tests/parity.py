"""Shared parity criteria (BASELINE.json north_star):
  * primary-hit primitive id bit-exact except at documented near-ties (the oracle's fragile mask, bit 0);
  * float buffer within 1e-4 relative error and the 8-bit image within 1 LSB on >= 99.9 % of pixels.
"""
import numpy as np

from oracle import oracle as O

REL_TOL = 1e-4          # north_star: float-buffer relative error <= 1e-4
GOOD_FRACTION = 0.999   # ... on at least 99.9 % of pixels


def oracle_rgb8(ref_rgb):
    rgb = ref_rgb.copy()
    O.normalize(rgb)
    return O.to_vec(rgb)


def check_exact(got, ref, rel=0.0):
    """FP64 validation mode: ids bit-exact everywhere, colours bit-exact up to `rel` (libm pow ulps)."""
    assert np.array_equal(got["prim_id"], ref["prim_id"])
    a, b = got["rgb"], ref["rgb"]
    if rel == 0.0:
        assert np.array_equal(a, b)
    else:
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        assert err.max() <= rel, err.max()


def check_fp32(got, ref, rows, rgb8=None):
    """Production FP32 mode against the f64 oracle over the rendered rows."""
    ids, rid = got["prim_id"][:rows], ref["prim_id"][:rows]
    mism = ids != rid
    fragile = (ref["fragile"][:rows] & 1) != 0
    assert not np.any(mism & ~fragile), "primitive id differs at %d pixels outside the near-tie mask" % int((mism & ~fragile).sum())
    a, b = got["rgb"][:rows].astype(np.float64), ref["rgb"][:rows]
    rel = (np.abs(a - b) / np.maximum(np.abs(b), 1e-12)).max(axis=2)
    frac_ok = float((rel <= REL_TOL).mean())
    assert frac_ok >= GOOD_FRACTION, "only %.5f of pixels within %g" % (frac_ok, REL_TOL)
    out = {"id_mismatch": int(mism.sum()), "fragile_frac": float(fragile.mean()), "frac_within_tol": frac_ok,
           "max_rel": float(rel.max())}
    if rgb8 is not None:
        d = np.abs(rgb8[:rows].astype(np.int16) - oracle_rgb8(ref["rgb"])[:rows].astype(np.int16)).max(axis=2)
        frac8 = float((d <= 1).mean())
        assert frac8 >= GOOD_FRACTION, "only %.5f of pixels within 1 LSB" % frac8
        out["frac_within_1lsb"] = frac8
        out["frac_exact_8bit"] = float((d == 0).mean())
    return out
