"""Host-side logic of the mirror package that needs no GPU: metric vectorisation, history flattening, configs."""
import numpy as np

from conftest import synthetic_data
from oracle import ref_host as H


def test_batch_metrics_bit_identical_to_reference_formula():
    from cleverrec_b200.utils.metrics import batch_ranking_metrics, cal_ranking_metrics
    rs = np.random.RandomState(1)
    for K in (1, 5, 10, 20):
        real_lists, recs = [], []
        for _ in range(300):
            real_lists.append(rs.permutation(50)[:rs.randint(1, 9)].tolist())
            recs.append(rs.permutation(50)[:20])
        recs = np.asarray(recs)
        recs[::7, 15:] = -1  # padded rows (fewer than K candidates)
        hr, mrr, ndcg = batch_ranking_metrics(real_lists, recs, K)
        for k in range(len(real_lists)):
            want = H.cal_ranking_metrics(real_lists[k], recs[k, :K], K)
            assert (hr[k], mrr[k], ndcg[k]) == want
            assert cal_ranking_metrics(real_lists[k], recs[k, :K], K) == want


def test_history_from_dict_matches_oracle_builder():
    from cleverrec_b200.engine import history_from_dict
    from oracle import philox as X
    d = synthetic_data(60, 150, 12, seed=4)
    d.ui_train[5] = d.ui_train[5] + d.ui_train[5][:3]  # duplicated interactions (Ciao has them)
    a = history_from_dict(d.ui_train, d.user_nums)
    b = X.build_history(d.ui_train, d.user_nums)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_initializers_and_config_parsing():
    import torch
    from cleverrec_b200.utils.tools import get_initializer, re_index
    g = torch.Generator().manual_seed(0)
    w = get_initializer('normal ', 0.01, g)([2000, 64])  # trailing blank as in conf/BPR.properties:13
    assert abs(float(w.std()) - 0.01) < 5e-4 and w.dtype == torch.float32
    x = get_initializer('xavier_uniform', 0.01, g)([100, 28])
    assert float(x.abs().max()) <= (6.0 / 128) ** 0.5 + 1e-7
    t = get_initializer('tnormal', 0.5, g)([4000])
    assert float(t.abs().max()) <= 1.0
    assert get_initializer('nope', 0.1) is None
    assert re_index(['b', 'a']) == {'b': 0, 'a': 1}


def test_checkpoint_helpers_round_trip(tmp_path):
    import time
    import torch
    from cleverrec_b200.utils.tools import latest_checkpoint, load_checkpoint, save_checkpoint
    assert latest_checkpoint(str(tmp_path / "missing")) is None and latest_checkpoint(None) is None
    a = save_checkpoint(str(tmp_path / "GMF"), "GMF", {"GMF_params/P": torch.arange(6.0).reshape(2, 3), "GMF_params/h_gmf": np.ones(3)})
    time.sleep(0.02)
    b = save_checkpoint(str(tmp_path / "GMF"), "GMF", {"GMF_params/P": torch.zeros(2, 3)}, step=7)
    assert latest_checkpoint(str(tmp_path / "GMF")) == b and a != b
    v = load_checkpoint(a)
    assert set(v) == {"GMF_params/P", "GMF_params/h_gmf"} and v["GMF_params/P"].dtype == np.float32
    assert v["GMF_params/P"].tolist() == [[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]]



def test_eval_cache_guard_drops_the_cached_item_table_on_torch_writes_and_reallocation():
    """Engine.score_topk's guard (no GPU needed: the library call is a stub).  The C side keys its bf16 copy of the item table by
    address; torch in-place writes (also through views) and a new tensor at a recycled address must drop it, repeated calls must not."""
    import torch
    from cleverrec_b200.engine import Engine

    class Lib(object):
        drops = 0

        def crb_eval_cache_invalidate(self, h):
            Lib.drops += 1
            return 0
    e = Engine.__new__(Engine)
    e.lib, e.h = Lib(), 1
    try:
        Q, hv = torch.zeros(10, 4), torch.zeros(4)
        steps = [
            (lambda: (Q, None), 1), (lambda: (Q, None), 1), (lambda: (Q[:5], None), 1),        # first use; same table; a view of it
            (lambda: (Q.mul_(2), None), 2), (lambda: (Q, None), 2),                          # in-place write
            (lambda: (Q[:3].add_(1)._base, None), 3),                                        # write through a view
            (lambda: (Q, hv), 4), (lambda: (Q, hv), 4), (lambda: (Q, hv.add_(1)), 5),        # hvec joins the key
            (lambda: (torch.zeros(10, 4), hv), 6), (lambda: (torch.zeros(10, 4), hv), 7),    # new tables (possibly at a recycled address)
        ]
        for make, want in steps:
            q, h = make()
            e._eval_cache_guard(q, h)
            assert Lib.drops == want
        e._eval_cache_guard(object(), None)      # not a tensor: dropped, not raised
        assert Lib.drops == 8
    finally:
        e.h = None
