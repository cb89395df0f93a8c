"""Host-side logic of the GP layer that needs no GPU: grouping a candidate list by covariance shape
(b200id.gp.shape_groups, the host half of b200id_gp_batch_eval_shared)."""
import numpy as np
import pytest

from oracle import gp as ogp


@pytest.fixture(scope="module")
def gp():
    from b200id import gp as mod
    return mod


@pytest.mark.parametrize("name", ["rbf", "matern12", "matern32", "matern52", "exp", "ss1", "ss2", "sshf", "stable_spline"])
def test_product_grid_groups_by_everything_but_the_amplitude(gp, name):
    """grid_search.py:173: the list is itertools.product of the per-parameter grids, so the 900 candidates of a
    two-parameter kernel are 30 shapes x 30 amplitudes; every candidate maps to a shape whose non-amplitude
    parameters are bitwise its own, and the first candidate of each shape is its representative."""
    cands = ogp.grid_candidates(name, 5000)
    kid = ogp.KERNELS[name][0]
    shape_of, shape_theta = gp.shape_groups(kid, cands)
    th = gp._theta_block(cands)
    assert shape_of.dtype == np.int32 and shape_of.shape == (len(cands),)
    assert shape_theta.shape == (30, gp.MAX_PARAMS)
    other = [c for c in range(gp.MAX_PARAMS) if c != gp.AMPLITUDE_SLOT]
    assert np.array_equal(shape_theta[shape_of][:, other].view(np.uint64), th[:, other].view(np.uint64))
    # the representative is a member of its own group
    for q, rep in enumerate(shape_theta):
        members = th[shape_of == q]
        assert len(members) == 30 and any(np.array_equal(rep, m) for m in members)


def test_lists_that_do_not_share_take_the_direct_path(gp, monkeypatch):
    rng = np.random.default_rng(3)
    assert gp.shape_groups(ogp.KERNELS["matern52"][0], rng.uniform(0.1, 2.0, size=(4096, 2))) is None      # restart points
    assert gp.shape_groups(ogp.KERNELS["di"][0], ogp.grid_candidates("di", 900)) is None                   # amplitude is theta[0]
    assert gp.shape_groups(ogp.KERNELS["rbf"][0], ogp.grid_candidates("rbf", 900)[:3]) is None             # too short to matter
    # DC (three parameters): the 27000-point product sampled down to 5000 still has ~5 amplitudes per (alpha, rho) pair,
    # sampled down to 900 it has about one
    dc = ogp.KERNELS["dc"][0]
    so, st = gp.shape_groups(dc, ogp.grid_candidates("dc", 5000))
    assert st.shape[0] * gp.SHARED_MIN_GAIN <= 5000 and so.max() == st.shape[0] - 1
    assert gp.shape_groups(dc, ogp.grid_candidates("dc", 900)) is None
    monkeypatch.setenv("B200ID_GP_SHARED", "0")
    assert gp.shape_groups(ogp.KERNELS["rbf"][0], ogp.grid_candidates("rbf", 900)) is None


def test_round_robin_shards_of_the_grid_still_share(gp):
    """Candidates are sharded round robin over the ranks; an 8-rank shard of the 900-point grid holds
    112-113 candidates over all 30 shapes, i.e. 3-4 amplitudes per shape: still above SHARED_MIN_GAIN."""
    cands = ogp.grid_candidates("matern52", 900)
    for world in (2, 4, 8):
        for rank in (0, world - 1):
            shard = cands[rank::world]
            groups = gp.shape_groups(ogp.KERNELS["matern52"][0], shard)
            assert groups is not None and groups[1].shape[0] == 30


@pytest.mark.parametrize("name", [k for k in ogp.KERNELS if k not in ("di", "matern_nu")])
def test_amplitude_is_the_last_factor_in_the_reference_arithmetic(name):
    """The basis of b200id_gp_batch_eval_shared, checked on the oracle's restatement of kernels.py:90-382 (pinned to
    reference-generated goldens): for every kernel but DI, K(theta) is bit for bit theta[1] * K(theta with theta[1] = 1)
    -- the amplitude multiplies last, 1 * s is exact, and (v * sign) * m == v * (sign * m) for sign = +-1 (SSHF)."""
    rng = np.random.default_rng(len(name))
    X = np.sort(rng.uniform(-1.7, 1.7, 40))
    Xv = rng.uniform(-1.6, 1.6, 25)
    cands = ogp.grid_candidates(name, 5000)
    for th in cands[rng.choice(len(cands), 40, replace=False)]:
        unit = np.array(th, dtype=np.float64); unit[1] = 1.0
        for a, b in ((X, None), (Xv, X)):
            with np.errstate(all="ignore"):
                K, S = ogp.gram(name, th, a, b), ogp.gram(name, unit, a, b)
                want = th[1] * S
            assert np.array_equal(np.isnan(K), np.isnan(want))
            ok = ~np.isnan(K)
            assert np.array_equal(K[ok].view(np.uint64), want[ok].view(np.uint64)), name
