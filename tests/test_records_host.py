"""The file-keyed record store -- host logic only, no GPU."""
import os
import time

import numpy as np
import pytest
from scipy.io import savemat


@pytest.fixture()
def store():
    from b200id import records
    records.STORE.clear()
    yield records.STORE
    records.STORE.clear()


def _write(path, n=200, seed=0):
    rng = np.random.default_rng(seed)
    t = np.arange(n) * 0.002
    y, u = rng.standard_normal(n), rng.standard_normal(n)
    savemat(str(path), {"output": np.vstack([t, y, u])})
    return t, u, y


def test_loaders_share_one_parse_and_hand_out_frozen_vectors(tmp_path, store):
    from src.frequency_transform.frf_estimator import load_time_u_y, describe_timebase
    from src.fir_model.fir_validation import load_mat_time_series
    from src.fir_model.fir_helpers import _load_timebase_from_mat
    t, u, y = _write(tmp_path / "a.mat")
    t1, u1, y1 = load_time_u_y(tmp_path / "a.mat")
    assert np.array_equal(t1, t) and np.array_equal(u1, u) and np.array_equal(y1, y)
    t2, u2, y2 = load_time_u_y(str(tmp_path / "a.mat"))
    assert t2 is t1 and u2 is u1 and y2 is y1                     # same objects: identity is the device-cache key
    assert store.stats["mat_misses"] == 1 and store.stats["record_hits"] == 1
    for a in (t1, u1, y1):
        assert not a.flags.writeable and a.flags.c_contiguous and a.dtype == np.float64
        with pytest.raises(ValueError):
            a[0] = 1.0
    ts = load_mat_time_series(tmp_path / "a.mat")
    assert ts["t"] is t1 and ts["u"] is u1 and ts["y"] is y1       # same vectors from both loaders: one shared record
    assert store.stats["record_aliases"] == 1
    tt, Ts = _load_timebase_from_mat(tmp_path / "a.mat")
    assert abs(Ts - 0.002) < 1e-15 and np.array_equal(tt, t)
    assert store.stats["mat_misses"] == 1, "three loaders, one loadmat"
    rec = store.lookup(t1, u1, y1)
    assert rec is not None and (rec.role_of(t1), rec.role_of(u1), rec.role_of(y1)) == ("t", "u", "y")
    assert store.lookup(t1, u1.copy()) is None and store.lookup(np.array(t1)) is None
    _write(tmp_path / "b.mat", seed=5)
    tb, ub, yb = load_time_u_y(tmp_path / "b.mat")
    assert store.lookup(t1, ub) is None and store.lookup(tb, ub, yb) is not rec   # records do not mix
    assert describe_timebase(t1) == describe_timebase(np.array(t1)) and rec.monotone() is True


def test_rewritten_file_misses(tmp_path, store):
    from src.frequency_transform.frf_estimator import load_time_u_y
    _write(tmp_path / "a.mat", seed=0)
    _, u_old, _ = load_time_u_y(tmp_path / "a.mat")
    time.sleep(0.01)
    _, u_new_expected, _ = _write(tmp_path / "a.mat", n=210, seed=1)
    os.utime(tmp_path / "a.mat", ns=(time.time_ns(), time.time_ns()))
    _, u_new, _ = load_time_u_y(tmp_path / "a.mat")
    assert u_new.size == 210 and np.array_equal(u_new, u_new_expected) and u_new is not u_old
    assert store.stats["mat_misses"] == 2


def test_lru_budget_and_disable(tmp_path, monkeypatch):
    from b200id import records
    st = records.RecordStore(host_budget=3 * 3 * 200 * 8 + 100, device_budget=0)
    paths = []
    for k in range(5):
        _write(tmp_path / f"r{k}.mat", seed=k)
        paths.append(tmp_path / f"r{k}.mat")
    ext = lambda d: (d["output"][0], d["output"][2], d["output"][1])
    recs = [st.record(p, "x", ext) for p in paths]
    assert len(st.records) <= 3 and (records.file_key(paths[-1]), "x") in st.records
    assert st.lookup(recs[0].t) is None and st.lookup(recs[-1].t) is recs[-1]      # evicted ids are forgotten
    assert sum(n for _, n in st.mats.values()) <= st.host_budget
    monkeypatch.setattr(records, "ENABLED", False)
    a = st.record(paths[0], "x", ext)
    b = st.record(paths[0], "x", ext)
    assert a is not b and st.lookup(a.t) is None
    st.clear()
    assert not st.records and not st.by_id and not st.mats


def test_extraction_errors_propagate(tmp_path, store):
    from src.frequency_transform.frf_estimator import load_time_u_y
    savemat(str(tmp_path / "bad.mat"), {"z": np.zeros((4, 4))})
    for _ in range(2):
        with pytest.raises(RuntimeError, match="Could not locate"):
            load_time_u_y(tmp_path / "bad.mat")
    assert not store.records and store.stats["mat_misses"] == 1
