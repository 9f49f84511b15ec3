"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU twin (NumPy, integer-exact) of the on-device triplet sampler in
``cleverrec_b200/csrc/sampler.cuh``.  The device sampler replaces the Python
loops of the reference samplers (utils/sampler.py:10-99) with a counter-based
scheme that gives the *same distribution* (uniform negatives outside the user's
history, distinct inside a positive's group, one global shuffle per epoch) but a
different random stream.  This twin pins the device output bit-for-bit.

Spec (shared by both implementations)
-------------------------------------
* Philox4x32-10 (Salmon et al. 2011), key = (seed_lo, seed_hi).
* Negatives of positive ``p`` in epoch ``e``: candidate words are the outputs of
  Philox blocks with counter (p_lo, p_hi, blk, e), blk = 0,1,2,..., words in order
  0..3.  A word w gives v = w & mask (mask = 2^ceil(log2 I) - 1); it is accepted iff
  v < I, v not in seen(u) and v not already accepted for this positive -- exactly
  the test at utils/sampler.py:58-61 with np.random.randint's masked rejection.
  Slot s of the positive is the s-th accepted value.
* Epoch shuffle (utils/sampler.py:68): position k holds source element pi(k), pi =
  6-round balanced Feistel network over 2^b >= N (b even) with cycle walking;
  round keys = Philox block with counter (0,0,0xFFFFFFFF-? ,e) -- see perm_keys().
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
U32 = np.uint64(0xFFFFFFFF)
MAX_BLOCKS = 4096  # attempt guard shared with the device code (CRB_SAMPLER_MAX_BLOCKS)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  All inputs broadcastable uint32-valued arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & U32
    c1 = np.asarray(c1, dtype=np.uint64) & U32
    c2 = np.asarray(c2, dtype=np.uint64) & U32
    c3 = np.asarray(c3, dtype=np.uint64) & U32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & U32
        hi1, lo1 = p1 >> np.uint64(32), p1 & U32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def perm_keys(seed, epoch):
    """Six Feistel round keys for the epoch shuffle (computed on the host in the product too)."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    a = philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0, epoch, k0, k1)
    b = philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 1, epoch, k0, k1)
    return [int(a[0]), int(a[1]), int(a[2]), int(a[3]), int(b[0]), int(b[1])]


def _mix(x, k):
    x = (x ^ np.uint64(k)) & U32
    x = (x * np.uint64(0x85EBCA6B)) & U32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & U32
    x ^= x >> np.uint64(16)
    return x


def perm_bits(n):
    b = 2
    while (1 << b) < n:
        b += 2
    return b


def feistel_perm(k, n, keys):
    """pi(k) for k in [0,n): array in, array out (uint64)."""
    k = np.asarray(k, dtype=np.uint64)
    b = perm_bits(n)
    h = np.uint64(b // 2)
    hm = np.uint64((1 << (b // 2)) - 1)
    x = k.copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        v = x[todo]
        l, r = v >> h, v & hm
        for key in keys:
            l, r = r, (l ^ (_mix(r, key) & hm))
        v = (l << h) | r
        x[todo] = v
        todo_idx = np.nonzero(todo)[0]
        todo[todo_idx[v < np.uint64(n)]] = False
    return x


def item_mask(item_nums):
    m = max(item_nums - 1, 0)
    for s in (1, 2, 4, 8, 16):
        m |= m >> s
    return m


def group_negatives(p_idx, users, seed, epoch, neg_ratio, item_nums, seen_rowptr, seen_cols):
    """Accepted negatives of the positives ``p_idx`` (global positive indices).  -> int32 [len(p_idx), neg_ratio]."""
    p_idx = np.asarray(p_idx, dtype=np.uint64)
    users = np.asarray(users)
    n = p_idx.shape[0]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    out = np.full((n, neg_ratio), -1, dtype=np.int64)
    cnt = np.zeros(n, dtype=np.int64)
    mask = item_mask(item_nums)
    rows = np.repeat(np.arange(seen_rowptr.shape[0] - 1, dtype=np.int64), np.diff(seen_rowptr))
    seen_keys = rows * item_nums + seen_cols.astype(np.int64)
    active = np.arange(n)
    blk = 0
    while active.size and blk < MAX_BLOCKS:
        pa = p_idx[active]
        words = philox4x32_10(pa & U32, pa >> np.uint64(32), blk, epoch, k0, k1)
        for w in words:
            v = (w.astype(np.int64)) & mask
            ok = (v < item_nums) & (cnt[active] < neg_ratio)
            # membership in the sorted history (global key = user * I + item is sorted because the CSR is)
            key = users[active].astype(np.int64) * item_nums + v
            pos = np.searchsorted(seen_keys, key)
            inb = pos < seen_keys.shape[0]
            seen = np.zeros(active.size, dtype=bool)
            seen[inb] = seen_keys[pos[inb]] == key[inb]
            ok &= ~seen
            dup = (out[active] == v[:, None]).any(axis=1)
            ok &= ~dup
            rows = active[ok]
            out[rows, cnt[rows]] = v[ok]
            cnt[rows] += 1
        active = active[cnt[active] < neg_ratio]
        blk += 1
    return out.astype(np.int32)


def sample_pairwise(seed, epoch, first, count, neg_ratio, item_nums, pos_user, pos_item, seen_rowptr, seen_cols):
    """Triplets at epoch positions [first, first+count) -> (u, i, j, nbr_num) int32 arrays.

    Twin of crb_sample_pairwise; restates utils/sampler.py:46-74 (same tuple, minus train_batches)."""
    n_pos = pos_user.shape[0]
    N = n_pos * neg_ratio
    keys = perm_keys(seed, epoch)
    k = np.arange(first, first + count, dtype=np.uint64)
    s = feistel_perm(k, N, keys)
    p = (s // np.uint64(neg_ratio)).astype(np.int64)
    slot = (s % np.uint64(neg_ratio)).astype(np.int64)
    u = pos_user[p]
    negs = group_negatives(p, u, seed, epoch, neg_ratio, item_nums, seen_rowptr, seen_cols)
    j = negs[np.arange(count), slot]
    nbr = (seen_rowptr[u + 1] - seen_rowptr[u]).astype(np.int32)
    return u.astype(np.int32), pos_item[p].astype(np.int32), j.astype(np.int32), nbr


def sample_pointwise(seed, epoch, first, count, neg_ratio, item_nums, pos_user, pos_item, seen_rowptr, seen_cols):
    """Rows at epoch positions [first, first+count) -> (u, i, y) ; restates utils/sampler.py:10-43."""
    n_pos = pos_user.shape[0]
    g = neg_ratio + 1
    N = n_pos * g
    keys = perm_keys(seed, epoch)
    k = np.arange(first, first + count, dtype=np.uint64)
    s = feistel_perm(k, N, keys)
    p = (s // np.uint64(g)).astype(np.int64)
    r = (s % np.uint64(g)).astype(np.int64)
    u = pos_user[p]
    negs = group_negatives(p, u, seed, epoch, neg_ratio, item_nums, seen_rowptr, seen_cols)
    it = np.where(r == 0, pos_item[p], negs[np.arange(count), np.maximum(r - 1, 0)])
    y = (r == 0).astype(np.float32)
    nbr = (seen_rowptr[u + 1] - seen_rowptr[u]).astype(np.int32)
    return u.astype(np.int32), it.astype(np.int32), y, nbr


def sample_cml(seed, epoch, first, count, neg_ratio, item_nums, pos_user, pos_item, seen_rowptr, seen_cols):
    """Rows at epoch positions [first, first+count) -> (u, i, neg[count,neg_ratio]); restates utils/sampler.py:77-99."""
    n_pos = pos_user.shape[0]
    keys = perm_keys(seed, epoch)
    k = np.arange(first, first + count, dtype=np.uint64)
    p = feistel_perm(k, n_pos, keys).astype(np.int64)
    u = pos_user[p]
    negs = group_negatives(p, u, seed, epoch, neg_ratio, item_nums, seen_rowptr, seen_cols)
    return u.astype(np.int32), pos_item[p].astype(np.int32), negs


def build_history(ui_train, user_nums):
    """dict[int -> list[int]] (reference data.ui_train) -> flat positives (reference enumeration order,
    utils/sampler.py:50-52) + sorted-unique CSR for membership tests."""
    pos_user, pos_item = [], []
    seen_rowptr = np.zeros(user_nums + 1, dtype=np.int64)
    cols = []
    per_user = {}
    for u, items in ui_train.items():
        pos_user.extend([u] * len(items))
        pos_item.extend(items)
        per_user[u] = np.unique(np.asarray(items, dtype=np.int64))
    for u in range(user_nums):
        if u in per_user:
            cols.append(per_user[u])
            seen_rowptr[u + 1] = seen_rowptr[u] + per_user[u].shape[0]
        else:
            seen_rowptr[u + 1] = seen_rowptr[u]
    seen_cols = np.concatenate(cols).astype(np.int32) if cols else np.zeros(0, np.int32)
    return (np.asarray(pos_user, dtype=np.int32), np.asarray(pos_item, dtype=np.int32), seen_rowptr, seen_cols)


def social_history(ui_train, user_friends, SPu, user_nums):
    """What the SBPR sampler walks (utils/sampler.py:105-131), flattened: positives of the users that have an SPu (dict order),
    per-user SPu lists with the social coefficient of every entry, and the sorted own + social item sets."""
    pos_user, pos_item = [], []
    spu, suk, excl = {}, {}, {}
    for u, items in ui_train.items():
        if u not in SPu:
            continue
        pos_user.extend([u] * len(items))
        pos_item.extend(items)
        spu[u] = list(SPu[u])
        suk[u] = [sum(1 for f in user_friends[u] if f in ui_train and k in ui_train[f]) for k in SPu[u]]   # :126-130
        excl[u] = sorted(set(items) | set(SPu[u]))
    return (np.asarray(pos_user, dtype=np.int32), np.asarray(pos_item, dtype=np.int32), spu, suk, excl)


def sample_sbpr(seed, epoch, first, count, neg_ratio, item_nums, social):
    """Rows at epoch positions [first, first+count) -> (u, i, k, j, suk).  Twin of crb_sample_sbpr: source row s = pi(position)
    draws from Philox blocks with counter (s_lo, s_hi, 0x80000000 | blk, epoch); the first word w with
    (w & mask(len(SPu[u]) - 1)) < len(SPu[u]) picks the social item, each later word is a negative candidate (masked rejection,
    then membership in own + social items).  Rows are independent: no distinctness inside a positive's group."""
    pos_user, pos_item, spu, suk, excl = social
    N = pos_user.shape[0] * neg_ratio
    keys = perm_keys(seed, epoch)
    src = feistel_perm(np.arange(first, first + count, dtype=np.uint64), N, keys)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    imask = item_mask(item_nums)
    out = np.zeros((5, count), dtype=np.int64)
    for t, s in enumerate(src.tolist()):
        p = s // neg_ratio
        u = int(pos_user[p])
        n_sp = len(spu[u])
        smask = item_mask(n_sp)
        ex = set(excl[u])
        pick, neg, blk = -1, -1, 0
        while neg < 0 and blk < MAX_BLOCKS:
            words = philox4x32_10(s & 0xFFFFFFFF, s >> 32, 0x80000000 | blk, epoch, k0, k1)
            for w in words:
                w = int(w)
                if neg >= 0:
                    break
                if pick < 0:
                    if (w & smask) < n_sp:
                        pick = w & smask
                    continue
                v = w & imask
                if v < item_nums and v not in ex:
                    neg = v
            blk += 1
        out[:, t] = (u, pos_item[p], spu[u][pick], neg, suk[u][pick])
    return tuple(out[r].astype(np.int32) for r in range(4)) + (out[4].astype(np.float32),)


def sample_eval_negatives(seed, test_users, neg_samples, item_nums, seen_rowptr, seen_cols):
    """Integer-exact twin of crb_prep_eval_negatives (csrc/preprocess.cu): per test user, neg_samples distinct unseen items.
    Restates the LAW of RankingPreprocess.py:120-129 (np.random.choice(list(item_set - seen), size, replace=False)): uniform without
    replacement over the unseen items.  Candidates come 32 per round -- lane l takes word l & 3 of Philox block
    (user, 0xE7A1, round * 8 + l // 4, 0xFFFFFFFD) under key (seed lo, seed hi) -- masked to the next power of two; a candidate is
    accepted in lane order when it is in range, unseen, not accepted before and the first of its value in the round."""
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    mask = item_mask(item_nums)
    out = np.full((len(test_users), neg_samples), -1, dtype=np.int32)
    lanes = np.arange(32)
    for k, u in enumerate(np.asarray(test_users).tolist()):
        seen = set(seen_cols[seen_rowptr[u]:seen_rowptr[u + 1]].tolist())
        got, taken, rnd = 0, set(), 0
        while got < neg_samples:
            assert rnd < 65536
            w = philox4x32_10(u, 0xE7A1, rnd * 8 + lanes // 4, 0xFFFFFFFD, k0, k1)
            words = np.stack(w, axis=1)[lanes, lanes & 3]
            in_round = set()
            for v in (words & np.uint32(mask)).tolist():
                if v < item_nums and v not in seen and v not in taken and v not in in_round:
                    in_round.add(v)
                    if got < neg_samples:
                        out[k, got] = v
                        taken.add(v)
                    got += 1
            rnd += 1
    return out
