"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

CPU restatement of the TensorFlow-1 graphs the reference builds for the hot path, plus the
TF-1 optimizer semantics they rely on.  PARITY UNPINNED for TensorFlow's own kernels
(summation order, fp32 rounding): TensorFlow "1.13+" (README.md:57) is an un-vendored
dependency that is not in /root/reference and cannot be installed here, and the reference
has no tests or golden values.  What pins this file: (i) the GENUINE
reference model classes are executed, unmodified, on a TF-1 API shim (oracle/tf1_shim.py) and
every function below reproduces their losses, updated variables and pre_scores to 1e-10 in
fp64 over three optimizer steps, for SGD / Adagrad / Adam, and over whole genuine epoch loops
(tests/test_reference_graphs.py) -- a mechanical check against the reference's own
graph-building lines; (ii) gradients come from *autograd* (an independent derivation from the
hand-written CUDA backward); (iii) fp64 finite-difference checks in
tests/test_oracle_restatement.py.

Forward graphs (each returns the scalar loss exactly as `self.loss`):
  bpr_loss     <- model/ranking/BPR.py:31-44
  mf_loss      <- (no source) BPR's dot product under get_loss('square'|'cross_entropy')
  gmf_loss     <- model/ranking/GMF.py:37-49
  neumf_loss   <- model/ranking/NeuMF.py:58-95 (MLP tower: MLP.py:44-53)
  mlp_loss     <- model/ranking/MLP.py:44-70
  cml_loss     <- model/ranking/CML.py:39-70
  fism_loss    <- model/ranking/FISM.py:40-63 + utils/tools.py:90-97
  nais_loss    <- model/ranking/NAIS_single.py:59-90
  transcf_loss <- model/ranking/TransCF.py:38-71 + utils/tools.py:100-113
  lrml_loss    <- model/ranking/LRML.py:42-64
  sbpr_loss    <- model/ranking/SBPR.py:38-57
Losses        <- utils/tools.py:66-76
Optimizers    <- utils/tools.py:79-87 with TF-1 defaults
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- losses
def get_loss(loss_func, y, logits=None, margin=None):
    if loss_func == "cross_entropy":  # tf.nn.sigmoid_cross_entropy_with_logits, summed
        x, z = logits, y
        return (torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x)))).sum()
    if loss_func == "bpr":  # -log_sigmoid(y) = softplus(-y)
        return F.softplus(-y).sum()
    if loss_func == "hinge":
        return torch.clamp(y + margin, min=0).sum()
    if loss_func == "square":
        return ((y - logits) ** 2).sum()
    raise ValueError(loss_func)


def l2_loss(t):  # tf.nn.l2_loss
    return 0.5 * (t * t).sum()


# ----------------------------------------------------------------------------- forward graphs
def bpr_loss(p, b, hp):
    ue, ie, je = p["P"][b["u"]], p["Q"][b["i"]], p["Q"][b["j"]]
    ui = (ue * ie).sum(1)
    uj = (ue * je).sum(1)
    return get_loss(hp.get("loss_func", "bpr"), ui - uj, margin=hp.get("margin")) + hp["reg"] * (l2_loss(ue) + l2_loss(ie) + l2_loss(je))


def mf_loss(p, b, hp):
    ue, ie = p["P"][b["u"]], p["Q"][b["i"]]
    logits = (ue * ie).sum(1)
    return get_loss(hp["loss_func"], b["y"], logits=logits) + hp["reg"] * (l2_loss(ue) + l2_loss(ie))


def gmf_loss(p, b, hp):
    ue, ie = p["P"][b["u"]], p["Q"][b["i"]]
    logits = ((ue * ie) * p["h"]).sum(1)
    return get_loss(hp["loss_func"], b["y"], logits=logits) + hp["reg"] * (l2_loss(ue) + l2_loss(ie))


def mlp_tower(x, p, n_layers):
    for k in range(n_layers):
        x = torch.relu(x @ p["W_%d" % k] + p["b_%d" % k])
    return x


def neumf_logits(p, u, i, n_layers):
    ug, ig = p["P_gmf"][u], p["Q_gmf"][i]
    um, im = p["P_mlp"][u], p["Q_mlp"][i]
    y_gmf = ug * ig
    y_mlp = mlp_tower(torch.cat([um, im], 1), p, n_layers)
    return (torch.cat([y_gmf, y_mlp], 1) * p["h_neumf"]).sum(1), (ug, ig, um, im)


def neumf_loss(p, b, hp):
    logits, (ug, ig, um, im) = neumf_logits(p, b["u"], b["i"], hp["n_layers"])
    return (get_loss(hp["loss_func"], b["y"], logits=logits) + hp["reg1"] * (l2_loss(ug) + l2_loss(ig))
            + hp["reg2"] * (l2_loss(um) + l2_loss(im)))


def mlp_loss(p, b, hp):
    """MLP.py:44-70: the tower alone -- concat(u, i) -> relu(x W_k + b_k) per layer -> . h_mlp; L2 on the two gathered embeddings."""
    um, im = p["P"][b["u"]], p["Q"][b["i"]]
    logits = mlp_tower(torch.cat([um, im], 1), p, hp["n_layers"]) @ p["h_mlp"]
    return get_loss(hp["loss_func"], b["y"], logits=logits) + hp["reg"] * (l2_loss(um) + l2_loss(im))


def cml_parts(p, b, hp):
    ue, ie = p["P"][b["u"]], p["Q"][b["i"]]
    ne = p["Q"][b["neg"]]  # [B, R, d]
    d_ui = ((ue - ie) ** 2).sum(1)
    d_un = ((ue[:, None, :] - ne) ** 2).sum(2)  # [B, R]
    d_min = d_un.min(1).values
    per_pair = torch.clamp(d_ui + hp["margin"] - d_min, min=0)
    imposters = (d_ui[:, None] + hp["margin"] - d_un) > 0
    rank = imposters.to(ue.dtype).mean(1) * hp["item_nums"] / hp["neg_ratio"]
    per_pair = per_pair * torch.log(rank + 1)
    return per_pair.sum()


def cml_cov_loss(p, hp):
    X = torch.cat((p["Q"], p["P"]), 0)
    n_rows = float(X.shape[0])
    X = X - X.mean(0)
    cov = (X.t() @ X) / n_rows
    cov = cov - torch.diag(torch.diag(cov))
    return hp["reg"] * cov.sum()


def cml_loss(p, b, hp):
    return cml_parts(p, b, hp) + cml_cov_loss(p, hp)


def fism_user_embed(p, b, hp, hist):
    """hist = (rows, cols, vals) COO of get_ui_sp_mat (vals 1/len(items), duplicates kept)."""
    rows, cols, vals = hist
    n_users = hp["user_nums"] + 1
    agg = torch.zeros(n_users, p["P"].shape[1], dtype=p["P"].dtype)
    agg = agg.index_add(0, rows, p["P"][cols] * vals[:, None])
    nbr = agg[b["u"]]
    coeff = torch.pow(b["nbr_num"].to(p["P"].dtype), -hp["alpha"])
    return coeff[:, None] * nbr


def fism_loss(p, b, hp, hist):
    s_u = fism_user_embed(p, b, hp, hist)
    ui = (p["Q"][b["i"]] * s_u).sum(1) + p["b"][b["i"]]
    loss = hp["reg"] * (l2_loss(p["P"]) + l2_loss(p["Q"])) / hp["batch_size"] + hp["reg_bias"] * l2_loss(p["b"])
    if "j" in b:
        uj = (p["Q"][b["j"]] * s_u).sum(1) + p["b"][b["j"]]
        return loss + get_loss(hp["loss_func"], ui - uj)
    return loss + get_loss(hp["loss_func"], b["y"], logits=ui)


def nais_user_embed(p, hist_items, q_rows, hp):
    ph = p["P"][hist_items]  # [n, d]
    if hp.get("atten_type", "prod") == "concat":
        T, n = q_rows.shape[0], ph.shape[0]
        joint = torch.cat([ph[None].expand(T, -1, -1), q_rows[:, None, :].expand(-1, n, -1)], 2)
    else:
        joint = q_rows[:, None, :] * ph[None, :, :]  # einsum('ac,bc->abc')
    a = torch.relu(joint @ p["W"] + p["b_att"]) @ p["h"]  # [T, n]
    e = torch.exp(a)
    d = torch.pow(e.sum(1, keepdim=True), hp["beta"])
    w = e / d
    return w @ ph  # [T, d]


def nais_loss(p, b, hp):
    q = p["Q"][b["i"]]
    bias = p["bias"][b["i"]]
    s = nais_user_embed(p, b["hist"], q, hp)
    x = (s * q).sum(1) + bias
    return get_loss("cross_entropy", b["y"], logits=x) + hp["reg"] * (l2_loss(s) + l2_loss(q) + l2_loss(bias))


def transcf_parts(p, b, hp, ui_coo, iu_coo):
    def spmm(coo, dense, n_rows):
        r, c, v = coo
        out = torch.zeros(n_rows, dense.shape[1], dtype=dense.dtype)
        return out.index_add(0, r, dense[c] * v[:, None])
    all_u = spmm(ui_coo, p["Q"], hp["user_nums"])
    all_i = spmm(iu_coo, p["P"], hp["item_nums"])
    return all_u, all_i


def transcf_loss(p, b, hp, ui_coo, iu_coo):
    all_u, all_i = transcf_parts(p, b, hp, ui_coo, iu_coo)
    ue, ie, je = p["P"][b["u"]], p["Q"][b["i"]], p["Q"][b["j"]]
    un, inb, jn = all_u[b["u"]], all_i[b["i"]], all_i[b["j"]]
    d_ui = ((ue + un * inb - ie) ** 2).sum(1)
    d_uj = ((ue + un * jn - je) ** 2).sum(1)
    loss = get_loss("hinge", d_ui - d_uj, margin=hp["margin"])
    reg_nbr = ((ue - un) ** 2).sum() + ((ie - inb) ** 2).sum()
    reg_dist = ((d_ui + hp["margin"] - d_uj) ** 2).sum()
    return loss + hp["reg1"] * reg_nbr + hp["reg2"] * reg_dist


def lrml_dist(p, ue, ie):
    """LRML.py:42-51 (_lram, training branch) + :59: relation vector from the memory module, translated squared distance."""
    joint = ue * ie
    att = torch.softmax(joint @ p["K"], dim=1)
    r = att @ p["M"]
    return ((ue + r - ie) ** 2).sum(1)


def lrml_loss(p, b, hp):
    ue, ie, je = p["P"][b["u"]], p["Q"][b["i"]], p["Q"][b["j"]]
    d_ui, d_uj = lrml_dist(p, ue, ie), lrml_dist(p, ue, je)
    return get_loss("hinge", d_ui - d_uj, margin=hp["margin"]) + hp["reg"] * (l2_loss(ue) + l2_loss(ie) + l2_loss(je))


def sbpr_loss(p, b, hp):
    """SBPR.py:38-57: u prefers its own item i over a friend's item k (scaled by 1/suk) and k over an unobserved j; the bias
    vector has item_nums + 1 entries (:36)."""
    ue = p["P"][b["u"]]
    def side(idx):
        e, bias = p["Q"][idx], p["bias"][idx]
        return e, bias, (ue * e).sum(1) + bias
    ie, ib, ui = side(b["i"])
    ke, kb, uk = side(b["k"])
    je, jb, uj = side(b["j"])
    return get_loss("bpr", (ui - uk) / b["suk"]) + get_loss("bpr", uk - uj) + hp["reg"] * (
        l2_loss(ue) + l2_loss(ie) + l2_loss(ke) + l2_loss(je) + l2_loss(ib) + l2_loss(kb) + l2_loss(jb))


# ----------------------------------------------------------------------------- scores (pre_scores)
def bpr_scores_pairs(P, Q, u, i):       # BPR.py:49
    return (P[u] * Q[i]).sum(1)


def bpr_scores_all(P, Q, users):        # BPR.py:51
    return P[users] @ Q.t()


def gmf_logits_pairs(P, Q, h, u, i):    # GMF.py:43 (ranking uses the logit; sigmoid is monotone)
    return ((P[u] * Q[i]) * h).sum(1)


def cml_dist_pairs(P, Q, u, i):         # CML.py:82
    return ((P[u] - Q[i]) ** 2).sum(1)


# ----------------------------------------------------------------------------- optimizers
class TF1Optimizer(object):
    """SGD / Adagrad / Adam with TF-1 default hyper-parameters and apply rules.

    step(params, grads, sparse_rows): ``sparse_rows[name]`` = 1-D LongTensor of *unique* touched rows for
    variables whose gradient is an IndexedSlices in TF (pure gathers); variables absent from it are dense.
    adam_mode: 'tf1' (moments decay for every row every step, all rows move) or 'lazy' (touched rows only).
    """

    def __init__(self, kind, lr, adam_mode="tf1"):
        self.kind, self.lr, self.adam_mode = kind, float(lr), adam_mode
        self.t = 0
        self.state = {}
        self.beta1, self.beta2, self.eps = 0.9, 0.999, 1e-8

    def _slot(self, name, like, init=0.0):
        if name not in self.state:
            self.state[name] = torch.full_like(like, init)
        return self.state[name]

    def lr_t(self):
        return self.lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)

    @torch.no_grad()
    def step(self, params, grads, sparse_rows=None):
        sparse_rows = sparse_rows or {}
        self.t += 1
        for name, var in params.items():
            g = grads.get(name)
            if g is None:
                continue
            rows = sparse_rows.get(name)
            if self.kind == "SGD":
                var -= self.lr * g
            elif self.kind == "Adagrad":
                acc = self._slot(name + "/acc", var, 0.1)
                if rows is None:
                    acc += g * g
                    var -= self.lr * g / torch.sqrt(acc)
                else:
                    acc[rows] += g[rows] * g[rows]
                    var[rows] -= self.lr * g[rows] / torch.sqrt(acc[rows])
            elif self.kind == "Adam":
                m, v = self._slot(name + "/m", var), self._slot(name + "/v", var)
                lr_t = torch.tensor(self.lr_t(), dtype=var.dtype)
                b1, b2 = torch.tensor(self.beta1, dtype=var.dtype), torch.tensor(self.beta2, dtype=var.dtype)
                eps = torch.tensor(self.eps, dtype=var.dtype)
                if rows is None:  # dense ApplyAdam
                    m += (g - m) * (1 - b1)
                    v += (g * g - v) * (1 - b2)
                    var -= (m * lr_t) / (torch.sqrt(v) + eps)
                elif self.adam_mode == "tf1":  # _apply_sparse_shared: dense decay + scatter_add + dense update
                    m *= b1
                    m[rows] += g[rows] * (1 - b1)
                    v *= b2
                    v[rows] += (g[rows] * g[rows]) * (1 - b2)
                    var -= lr_t * m / (torch.sqrt(v) + eps)
                else:  # lazy: LazyAdamOptimizer semantics, documented deviation
                    m[rows] = m[rows] * b1 + g[rows] * (1 - b1)
                    v[rows] = v[rows] * b2 + (g[rows] * g[rows]) * (1 - b2)
                    var[rows] -= lr_t * m[rows] / (torch.sqrt(v[rows]) + eps)
            else:
                raise ValueError(self.kind)


def train_step(loss_fn, params, batch, hp, opt, sparse_index=None, extra=()):
    """One `sess.run([train, loss])`.  ``sparse_index``: {var name: [batch keys whose values index it]} for
    variables reached only through gathers (IndexedSlices); everything else gets a dense apply."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    loss = loss_fn(leaves, batch, hp, *extra)
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    gdict = {k: g for k, g in zip(leaves.keys(), grads) if g is not None}
    rows = {}
    for name, keys in (sparse_index or {}).items():
        idx = torch.cat([batch[k].reshape(-1) for k in keys])
        rows[name] = torch.unique(idx)
    opt.step(params, gdict, rows)
    return float(loss.detach())


def to_torch(d, dtype=torch.float32):
    out = {}
    for k, v in d.items():
        a = np.asarray(v)
        if a.dtype.kind == "f":
            out[k] = torch.tensor(a, dtype=dtype)
        else:
            out[k] = torch.tensor(a.astype(np.int64))
    return out


def bpr_step_rowsparse(params, batch, reg, opt):
    """BPR step with IndexedSlices-style handling (gradients only for gathered rows, de-duplicated by unique +
    segment-sum like TF's _deduplicate_indexed_slices) and a row-sparse Adam/Adagrad/SGD apply.  Same maths as
    train_step(bpr_loss, ...) with adam_mode='lazy' but without materialising dense table gradients -- this is the
    version bench.py times as the CPU baseline (the generous one for the CPU)."""
    P, Q = params["P"], params["Q"]
    u, i, j = batch["u"], batch["i"], batch["j"]
    ue, ie, je = (t.detach().requires_grad_(True) for t in (P[u], Q[i], Q[j]))
    x = (ue * ie).sum(1) - (ue * je).sum(1)
    loss = F.softplus(-x).sum() + reg * (l2_loss(ue) + l2_loss(ie) + l2_loss(je))
    gu, gi, gj = torch.autograd.grad(loss, [ue, ie, je])
    opt.t += 1
    with torch.no_grad():
        for name, idx, g in (("P", u, gu), ("Q", torch.cat([i, j]), torch.cat([gi, gj]))):
            rows, inv = torch.unique(idx, return_inverse=True)
            G = torch.zeros(rows.shape[0], g.shape[1], dtype=g.dtype).index_add_(0, inv, g)
            var = params[name]
            if opt.kind == "SGD":
                var[rows] -= opt.lr * G
            elif opt.kind == "Adagrad":
                acc = opt._slot(name + "/acc", var, 0.1)
                a = acc[rows] + G * G
                acc[rows] = a
                var[rows] -= opt.lr * G / torch.sqrt(a)
            else:
                m, v = opt._slot(name + "/m", var), opt._slot(name + "/v", var)
                mr = m[rows] * opt.beta1 + G * (1 - opt.beta1)
                vr = v[rows] * opt.beta2 + (G * G) * (1 - opt.beta2)
                m[rows], v[rows] = mr, vr
                var[rows] -= opt.lr_t() * mr / (torch.sqrt(vr) + opt.eps)
    return float(loss.detach())
