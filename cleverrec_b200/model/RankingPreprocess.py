# coding: utf-8
""" Load and split a ranking dataset -- mirror of the reference model/RankingPreprocess.py:12-134.

Same constructor `(configs, logger)`, same config keys (`data.root_dir data.dataset data.file_name data.sep data.format
data.user_min data.item_min data.split_way data.split_by_time data.split_ratio test.neg_samples`), same attributes
(`user_nums item_nums ui_train ui_test`), and -- under the same `np.random.seed` -- the same split and the same sampled
evaluation negatives, bit for bit (tests/test_preprocess.py against golden vectors made by the genuine reference class):
the reference's results depend on pandas' groupby order, on sklearn's `train_test_split` permutation and on the iteration
order of Python sets of ids, so those three are used here exactly where the reference uses them; everything per-row is
vectorised.

Additions for the B200 path: `train_rows = (users int32[n], items int32[n])`, the training split as two columns in the
order the dict enumerates it, so that a model can hand them to `Engine.build_history` (csrc/history.cu) instead of walking
the dict; `lazy_dicts=True` skips building `ui_train` as a dict of Python lists when only the columns are needed.
`social_file` (RankingPreprocess.py:49-58) gives `user_friends: dict[int -> list[int]]` for SBPR; the SAMN padding (:60-67) belongs to
an out-of-scope model and is not applied."""
import os

import numpy as np
import pandas as pd


class RankingPreprocess(object):
    def __init__(self, configs, logger, lazy_dicts=False):
        self.configs, self.logger = configs, logger
        self.file_path = os.path.join(configs['data.root_dir'], configs['data.dataset'])
        ratings, item_set = self._load_data()
        self.ui_train, self.ui_test = self._split_data(ratings, item_set, lazy_dicts)

    # ------------------------------------------------------------------------------------------ load, filter, re-index
    def _load_data(self):
        c = self.configs
        fmt = c['data.format']
        names = {'UI': ['u_id', 'i_id'], 'UIR': ['u_id', 'i_id', 'rating'], 'UIRT': ['u_id', 'i_id', 'rating', 'time']}[fmt]
        # header=0 as in the reference (RankingPreprocess.py:22-32): the first line of the file is consumed as a header
        # even for header-less files such as ml-100k's u.data (SURVEY 2.3 'header quirk') -- kept, it changes the data
        ratings = pd.read_csv(os.path.join(self.file_path, c['data.file_name']), sep=c['data.sep'], header=0, names=names,
                              usecols=list(range(len(names))))
        if fmt == 'UIRT':
            ratings['time'] = ratings['time'].astype(int)
        user_min, item_min = int(c['data.user_min']), int(c['data.item_min'])
        if user_min > 0:   # users first, then items counted on what is left (RankingPreprocess.py:35-39)
            ratings = self._drop_rare(ratings, 'u_id', user_min)
        if item_min > 0:
            ratings = self._drop_rare(ratings, 'i_id', item_min)
        # New ids follow the iteration order of a Python set of the raw ids (utils/tools.py:9-15 re_index over a set)
        user_ids, item_ids = set(ratings['u_id'].unique()), set(ratings['i_id'].unique())
        self.user_nums, self.item_nums = len(user_ids), len(item_ids)
        raw_users = ratings['u_id'].to_numpy()
        ratings['u_id'] = self._renumber(raw_users, user_ids)
        ratings['i_id'] = self._renumber(ratings['i_id'].to_numpy(), item_ids)
        if 'social_file' in c:   # RankingPreprocess.py:49-58: trust pairs among the kept users, re-indexed, grouped by the trustor
            trusts = pd.read_csv(os.path.join(self.file_path, c['social_file']), sep=c['data.sep'], header=0, names=['u_id', 'v_id'], usecols=[0, 1])
            tu, tv = trusts['u_id'].to_numpy(), trusts['v_id'].to_numpy()
            known = np.fromiter(user_ids, dtype=raw_users.dtype, count=len(user_ids))
            keep = np.isin(tu, known) & np.isin(tv, known)
            tu, tv = self._renumber(tu[keep].astype(raw_users.dtype), user_ids), self._renumber(tv[keep].astype(raw_users.dtype), user_ids)
            order = np.argsort(tu, kind='stable')        # groupby('u_id').v_id.apply(list): ascending trustor, file order inside
            self.user_friends = _to_dict(tu[order], tv[order])
        return ratings, set(ratings['i_id'].unique())

    @staticmethod
    def _drop_rare(ratings, column, minimum):
        counts = ratings[column].map(ratings[column].value_counts())
        return ratings[counts >= minimum].reset_index(drop=True)

    @staticmethod
    def _renumber(raw, id_set):
        order = np.fromiter(id_set, dtype=raw.dtype, count=len(id_set))   # position in the set's iteration = new id
        sorter = np.argsort(order, kind='stable')
        return sorter[np.searchsorted(order, raw, sorter=sorter)]

    # ------------------------------------------------------------------------------------------ split
    def _split_data(self, ratings, item_set, lazy_dicts=False):
        c = self.configs
        split_way = c['data.split_way']
        if c['data.split_by_time'] == 'True':
            ratings.sort_values(['u_id', 'time'], inplace=True)   # same call, same (unstable quicksort) tie order
        if split_way == 'loo':
            # per user, in groupby (ascending id) order with the frame's row order inside a user: users with <= 3 rows are all
            # training, the others keep their last row for testing (RankingPreprocess.py:99-109)
            u = ratings['u_id'].to_numpy()
            order = np.argsort(u, kind='stable')
            us = u[order]
            last = np.r_[us[1:] != us[:-1], True] if len(us) else np.zeros(0, dtype=bool)
            size = np.bincount(us, minlength=self.user_nums)[us]
            is_test = last & (size > 3)
            train_data, test_data = ratings.iloc[order[~is_test]], ratings.iloc[order[is_test]]
        else:
            from sklearn.model_selection import train_test_split
            r1, r2, r3 = tuple(map(float, c['data.split_ratio'][1:-1].split(',')))
            if r2 > 0:
                train_data, rest = train_test_split(ratings, test_size=1.0 - r1)
                _, test_data = train_test_split(rest, test_size=r3 / (r2 + r3))
            else:
                train_data, test_data = train_test_split(ratings, test_size=r3)
        tu, ti = self._grouped_columns(train_data)
        self.train_rows = (tu.astype(np.int32), ti.astype(np.int32))
        ui_train = _LazyDict(tu, ti) if lazy_dicts else _to_dict(tu, ti)
        ui_test = _to_dict(*self._grouped_columns(test_data))
        # 99 / 1000 sampled negatives per test user, ground truth appended after them (RankingPreprocess.py:120-129)
        neg_samples = int(c['test.neg_samples'])
        if split_way == 'loo' or neg_samples > 0:
            with_negatives = {}
            for user in ui_test:
                seen = set(ui_train[user]) if user in ui_train else set()
                # np.random.choice over list(set difference): the candidate ORDER (Python set iteration) decides which ids the
                # drawn positions name, so the same set expression is evaluated
                drawn = np.random.choice(list(item_set - seen), size=neg_samples, replace=False).tolist()
                drawn.extend(ui_test[user])
                with_negatives[user] = drawn
            ui_test = with_negatives
        ratio = '' if split_way == 'loo' else ('split_ratio=%s, ' % c['data.split_ratio'])
        self.logger.info(' Data: dataset=%s, split_way=%s, neg_samples=%d, %suser_nums=%d, item_nums=%d, ratings_num=%d' % (
            c['data.dataset'], split_way, neg_samples, ratio, self.user_nums, self.item_nums, ratings.shape[0]))
        return ui_train, ui_test

    @staticmethod
    def _grouped_columns(frame):
        """(users, items) of a frame in `groupby('u_id').i_id.apply(list)` enumeration order: ascending user, frame order inside."""
        u, i = frame['u_id'].to_numpy(), frame['i_id'].to_numpy()
        order = np.argsort(u, kind='stable')
        return u[order], i[order]


def _to_dict(users, items):
    """dict user -> list of items, keys ascending (what `.groupby('u_id').i_id.apply(list).to_dict()` returns)."""
    if len(users) == 0:
        return {}
    cuts = np.flatnonzero(users[1:] != users[:-1]) + 1
    keys = users[np.r_[0, cuts]]
    return {int(k): part.tolist() for k, part in zip(keys, np.split(items, cuts))}


class _LazyDict(dict):
    """ui_train for very large splits: behaves like the dict of lists (keys, `in`, `[]`, len, iteration order) but keeps the two
    sorted columns and cuts a list out only when one is asked for."""

    def __init__(self, users, items):
        dict.__init__(self)
        self._items = items
        cuts = np.flatnonzero(users[1:] != users[:-1]) + 1 if len(users) else np.zeros(0, dtype=np.int64)
        self._keys = users[np.r_[0, cuts]] if len(users) else users
        self._start = np.r_[0, cuts, len(users)]
        self._pos = {int(k): n for n, k in enumerate(self._keys)}

    def __contains__(self, k): return k in self._pos
    def __len__(self): return len(self._pos)
    def __iter__(self): return iter(self._pos)
    def keys(self): return self._pos.keys()
    def __getitem__(self, k):
        n = self._pos[k]
        return self._items[self._start[n]:self._start[n + 1]].tolist()
    def get(self, k, default=None): return self[k] if k in self._pos else default
    def items(self): return ((k, self[k]) for k in self._pos)
    def values(self): return (self[k] for k in self._pos)


class DeviceRankingPreprocess(object):
    """`data.preprocess=device`: the same constructor and attributes, with filter / re-index / leave-one-out split / evaluation
    negatives on the device (csrc/preprocess.cu) and the history built natively from the training columns (csrc/history.cu) --
    no frame-sized Python object is ever built.  The file is still parsed on the host (I/O), then lives as device columns.

    What is the reference's bit for bit: user / item counts, the re-indexing (dense non-negative raw ids: the ascending order a
    Python set of them iterates in), the leave-one-out split, and -- `data.split_way=rs` -- the random split, whose permutation
    is drawn by the reference's own call (sklearn `train_test_split` on NumPy's global stream) and only APPLIED on the device.
    What is not NumPy's stream: the sampled evaluation negatives (`np.random.choice(list(item_set - seen), ...)` permutes the whole
    unseen catalogue per user, O(users x items)); they are drawn by crb_prep_eval_negatives -- same law, keyed by `seed` -- so
    HR / NDCG on sampled candidates match the packaged host mode in distribution, not per draw.  `social_file` is not supported here."""

    def __init__(self, configs, logger, engine=None):
        import torch
        from ..engine import Engine
        self.configs, self.logger = configs, logger
        c = configs
        if 'social_file' in c:
            raise NotImplementedError('data.preprocess=device does not load social_file; use the packaged host preprocessing')
        self.file_path = os.path.join(c['data.root_dir'], c['data.dataset'])
        self.engine = engine if engine is not None else Engine(int(c.get('engine.device', 0)))
        eng = self.engine
        fmt = c['data.format']
        names = {'UI': ['u_id', 'i_id'], 'UIR': ['u_id', 'i_id', 'rating'], 'UIRT': ['u_id', 'i_id', 'rating', 'time']}[fmt]
        frame = pd.read_csv(os.path.join(self.file_path, c['data.file_name']), sep=c['data.sep'], header=0, names=names, usecols=list(range(len(names))))
        res = eng.prep_filter_reindex(frame['u_id'].to_numpy(), frame['i_id'].to_numpy(), int(c['data.user_min']), int(c['data.item_min']))
        self.user_nums, self.item_nums = res['n_users'], res['n_items']
        u, i = res['u'], res['i']
        n = int(u.numel())
        split_way = c['data.split_way']
        if split_way == 'loo':
            time = None
            if c['data.split_by_time'] == 'True':
                time = torch.from_numpy(frame['time'].astype(int).to_numpy()).to(eng.device)[res['row']]
            perm, is_test = eng.prep_split_loo(u, self.user_nums, time)
            tr, te = perm[~is_test], perm[is_test]
        else:
            from sklearn.model_selection import train_test_split
            r1, r2, r3 = tuple(map(float, c['data.split_ratio'][1:-1].split(',')))
            if c['data.split_by_time'] == 'True':   # the reference sorts the frame by (user, time) before it permutes the rows
                time = torch.from_numpy(frame['time'].astype(int).to_numpy()).to(eng.device)[res['row']]
                perm, _ = eng.prep_split_loo(u, self.user_nums, time)
                u, i = u[perm].contiguous(), i[perm].contiguous()
            rows = np.arange(n)
            if r2 > 0:
                tr, rest = train_test_split(rows, test_size=1.0 - r1)
                _, te = train_test_split(rest, test_size=r3 / (r2 + r3))
            else:
                tr, te = train_test_split(rows, test_size=r3)
            tr, te = torch.from_numpy(tr).to(eng.device), torch.from_numpy(te).to(eng.device)
        # the native history builder groups the training rows by user (stable): its pos_user / pos_item ARE the dict's enumeration
        hist = eng.build_history(u[tr], i[tr], self.user_nums, self.item_nums)
        tu, ti = hist[0].cpu().numpy(), hist[1].cpu().numpy()
        self.train_rows = (tu, ti)
        self.ui_train = _LazyDict(tu, ti)
        # test rows grouped by user, frame order inside (groupby('u_id').i_id.apply(list))
        t_u, t_i = u[te], i[te]
        order = torch.sort(t_u.long(), stable=True)[1]
        t_u, t_i = t_u[order].cpu().numpy(), t_i[order].cpu().numpy()
        ui_test = _to_dict(t_u, t_i)
        neg_samples = int(c['test.neg_samples'])
        if split_way == 'loo' or neg_samples > 0:
            users = np.fromiter(ui_test.keys(), dtype=np.int32, count=len(ui_test))
            negs = eng.prep_eval_negatives(int(c.get('seed', 0)), users, neg_samples).cpu().numpy() if neg_samples > 0 else np.zeros((len(users), 0), np.int32)
            ui_test = {int(k): negs[pos].tolist() + ui_test[int(k)] for pos, k in enumerate(users)}
        self.ui_test = ui_test
        ratio = '' if split_way == 'loo' else ('split_ratio=%s, ' % c['data.split_ratio'])
        logger.info(' Data: dataset=%s, split_way=%s, neg_samples=%d, %suser_nums=%d, item_nums=%d, ratings_num=%d' % (
            c['data.dataset'], split_way, neg_samples, ratio, self.user_nums, self.item_nums, n))
