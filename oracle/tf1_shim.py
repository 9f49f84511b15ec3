"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal stand-in for the TensorFlow-1 graph API, just wide enough to EXECUTE the reference's own, unmodified model classes
(`/root/reference/model/ranking/{BPR,GMF,MLP,NeuMF,CML,FISM,NAIS_single,TransCF,LRML,SBPR}.py`, `model/Recommender.py`,
`utils/tools.py`) in the build container, where TensorFlow cannot be installed.  Installed as `sys.modules['tensorflow']` by
`load_reference_models()`, it lets `build_model()` build the reference's graph and `sess.run([self.train, self.loss], feed_dict)`
run it: every graph tensor is a lazy node over torch (fp64 by default), gradients come from torch.autograd, the three optimizers
follow TF-1's documented update rules (tf.train.GradientDescentOptimizer / AdagradOptimizer(initial_accumulator_value=0.1) /
AdamOptimizer(beta1=0.9, beta2=0.999, epsilon=1e-8) with lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) and epsilon OUTSIDE the
square root).  For the variables the models reach through gathers TF-1 builds IndexedSlices gradients and applies them with
`_apply_sparse` after summing duplicate indices; for these three optimizers that is the dense update on the densified gradient
(Adam: the moments of EVERY row decay and every row moves -- SURVEY.md 2.4), which is what is applied here.

What this pins and what it does not.  tests/test_reference_graphs.py runs the genuine reference graph code through this shim and
compares losses, updated variables and `pre_scores` with oracle/tf1_restatement.py in fp64 (1e-10): the restatement is thereby
checked MECHANICALLY against the reference's own graph-building lines instead of by reading them.  It does not pin TensorFlow's
kernels (summation order, fused ops, fp32 rounding): those stay unpinned, as DESIGN.md section 4 says.

Op semantics follow the TF-1 API documentation: tf.nn.l2_loss = sum(t^2) / 2; tf.clip_by_norm(t, c, axes) = t * c / max(||t||, c);
tf.nn.sigmoid_cross_entropy_with_logits = max(x, 0) - x z + log(1 + exp(-|x|)); tf.log_sigmoid(x) = -softplus(-x);
tf.reduce_min's gradient goes to the minimum (ties split evenly); tf.matrix_set_diag replaces the main diagonal."""
import contextlib
import importlib
import math
import sys
import types

import numpy as np
import torch

FLOAT = torch.float64          # the dtype every tf.float16 / tf.float32 tensor is computed in
_GEN = torch.Generator().manual_seed(0)
_VARIABLES = []

float16, float32, float64, int32, int64 = "float16", "float32", "float64", "int32", "int64"


def _torch_dtype(dt):
    if dt is None:
        return None
    return torch.int64 if str(dt).startswith("int") else (torch.bool if str(dt) == "bool" else FLOAT)


class _Env(object):
    def __init__(self, feeds):
        self.feeds, self.cache = feeds, {}


class Tensor(object):
    """A node of the lazy graph.  Hashable by identity (feed_dict keys), arithmetic operators build new nodes."""

    def __init__(self, fn, name=None):
        self._fn, self.name = fn, name

    def eval_in(self, env):
        k = id(self)
        if k not in env.cache:
            env.cache[k] = self._fn(env)
        return env.cache[k]

    def _bin(self, other, f, swap=False):
        return Tensor(lambda env: f(_val(other, env), self.eval_in(env)) if swap else f(self.eval_in(env), _val(other, env)))

    def __add__(self, o): return self._bin(o, lambda a, b: a + b)
    def __radd__(self, o): return self._bin(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._bin(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._bin(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b)
    def __rtruediv__(self, o): return self._bin(o, lambda a, b: a / b, True)
    def __neg__(self): return Tensor(lambda env: -self.eval_in(env))
    def __gt__(self, o): return self._bin(o, lambda a, b: a > b)
    def __lt__(self, o): return self._bin(o, lambda a, b: a < b)
    def __ge__(self, o): return self._bin(o, lambda a, b: a >= b)
    def __le__(self, o): return self._bin(o, lambda a, b: a <= b)
    def __getitem__(self, idx): return Tensor(lambda env: self.eval_in(env)[idx])
    __hash__ = object.__hash__


def _val(x, env):
    """Graph node -> its value in this run; Python / NumPy data -> a torch constant (floats in FLOAT, integers in int64)."""
    if isinstance(x, Tensor):
        return x.eval_in(env)
    if isinstance(x, torch.Tensor):
        return x
    a = np.asarray(x)
    if a.dtype.kind == "f":
        return torch.as_tensor(a, dtype=FLOAT)
    if a.dtype.kind == "b":
        return torch.as_tensor(a)
    return torch.as_tensor(a.astype(np.int64))


def _op(f, *args):
    return Tensor(lambda env: f(*[_val(a, env) for a in args]))


def _int(x, env):
    v = _val(x, env)
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)


# ------------------------------------------------------------------------------------------------ inputs, variables
class _Placeholder(Tensor):
    def __init__(self, dtype, shape, name):
        self.dtype = dtype
        Tensor.__init__(self, self._read, name)

    def _read(self, env):
        if id(self) not in env.feeds:
            raise KeyError("placeholder %r was not fed" % self.name)
        a = np.asarray(env.feeds[id(self)])
        return torch.as_tensor(a.astype(np.int64) if str(self.dtype).startswith("int") else a.astype(np.float64)).to(_torch_dtype(self.dtype))


def placeholder(dtype, shape=None, name=None):
    return _Placeholder(dtype, shape, name)


class Variable(Tensor):
    def __init__(self, initial_value, name=None, dtype=None, trainable=True):
        v = _val(initial_value, _Env({}))
        self.value = v.detach().clone().to(FLOAT).requires_grad_(True)
        Tensor.__init__(self, lambda env: self.value, name)
        _VARIABLES.append(self)

    def assign_value(self, array):
        with torch.no_grad():
            self.value.copy_(torch.as_tensor(np.asarray(array), dtype=FLOAT))

    def numpy(self):
        return self.value.detach().numpy().copy()


def get_variable(name, shape=None, dtype=None, initializer=None, regularizer=None, trainable=True):
    if callable(initializer) and not isinstance(initializer, Tensor):
        initializer = initializer(shape)
    return Variable(initializer, name=name)


def global_variables_initializer():
    return Tensor(lambda env: None, "init")


def reset_default_graph():
    del _VARIABLES[:]


def seed_initializers(seed):
    _GEN.manual_seed(seed)


@contextlib.contextmanager
def name_scope(name, *a, **k):
    yield name


variable_scope = name_scope


# ------------------------------------------------------------------------------------------------ initializers, constants
def random_normal_initializer(mean=0.0, stddev=1.0, **k):
    return lambda shape, **kk: torch.randn(*shape, generator=_GEN, dtype=FLOAT) * stddev + mean


def truncated_normal_initializer(mean=0.0, stddev=1.0, **k):
    def init(shape, **kk):
        t = torch.empty(*shape, dtype=FLOAT)
        torch.nn.init.trunc_normal_(t, mean=mean, std=stddev, a=mean - 2 * stddev, b=mean + 2 * stddev, generator=_GEN)
        return t
    return init


def random_uniform_initializer(minval=0.0, maxval=1.0, **k):
    return lambda shape, **kk: torch.rand(*shape, generator=_GEN, dtype=FLOAT) * (maxval - minval) + minval


def _xavier_initializer(uniform=True, **k):
    def init(shape, **kk):
        fi, fo = (shape[0], shape[0]) if len(shape) == 1 else (shape[0], shape[1])
        if uniform:
            lim = math.sqrt(6.0 / (fi + fo))
            return (torch.rand(*shape, generator=_GEN, dtype=FLOAT) * 2 - 1) * lim
        sd = math.sqrt(1.3 * 2.0 / (fi + fo))
        t = torch.empty(*shape, dtype=FLOAT)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=sd, a=-2 * sd, b=2 * sd, generator=_GEN)
        return t
    return init


def random_uniform(shape, minval=0.0, maxval=1.0, dtype=None, **k):
    if isinstance(shape, int):
        shape = [shape]
    v = torch.rand(*shape, generator=_GEN, dtype=FLOAT) * (maxval - minval) + minval
    return Tensor(lambda env: v)


def zeros(shape, dtype=float32, name=None):
    shape = [shape] if isinstance(shape, int) else list(shape)
    return Tensor(lambda env: torch.zeros(*shape, dtype=_torch_dtype(dtype)))


def constant(value, dtype=None, **k):
    return Tensor(lambda env: _val(value, env) if dtype is None else _val(value, env).to(_torch_dtype(dtype)))


class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        idx = np.asarray(indices, dtype=np.int64).reshape(-1, 2)
        self.rows, self.cols = torch.as_tensor(idx[:, 0].copy()), torch.as_tensor(idx[:, 1].copy())
        self.values = torch.as_tensor(np.asarray(values, dtype=np.float64), dtype=FLOAT)
        self.dense_shape = tuple(int(s) for s in dense_shape)


def sparse_tensor_dense_matmul(sp, dense, **k):
    def f(d):
        out = torch.zeros(sp.dense_shape[0], d.shape[1], dtype=d.dtype)
        return out.index_add(0, sp.rows, d[sp.cols] * sp.values[:, None])
    return _op(f, dense)


# ------------------------------------------------------------------------------------------------ ops
def _axis(axis):
    if axis is None:
        return None
    return tuple(axis) if isinstance(axis, (list, tuple)) else int(axis)


def _reduce(kind):
    def op(t, axis=None, keepdims=False, keep_dims=None, name=None, reduction_indices=None):
        keep = bool(keep_dims) if keep_dims is not None else bool(keepdims)
        ax = _axis(axis if axis is not None else reduction_indices)

        def f(x):
            if kind == "sum":
                return x.sum() if ax is None else x.sum(dim=ax, keepdim=keep)
            if kind == "mean":
                return x.mean() if ax is None else x.mean(dim=ax, keepdim=keep)
            return x.amin() if ax is None else x.amin(dim=ax, keepdim=keep)   # amin: ties share the gradient, as in TF
        return _op(f, t)
    return op


reduce_sum, reduce_mean, reduce_min = _reduce("sum"), _reduce("mean"), _reduce("min")


def einsum(equation, *operands):
    return _op(lambda *xs: torch.einsum(equation, *xs), *operands)


def gather(params, indices, **k):
    return _op(lambda p, i: p[i.long()], params, indices)


def expand_dims(t, axis=None, dim=None, name=None):
    ax = axis if axis is not None else dim
    return _op(lambda x: x.unsqueeze(ax), t)


def clip_by_norm(t, clip_norm, axes=None, name=None):
    ax = _axis(axes)

    def f(x):
        n = torch.sqrt((x * x).sum() if ax is None else (x * x).sum(dim=ax, keepdim=True))
        return x * clip_norm / torch.clamp(n, min=clip_norm)
    return _op(f, t)


def square(t, name=None): return _op(lambda x: x * x, t)
def squared_difference(a, b, name=None): return _op(lambda x, y: (x - y) * (x - y), a, b)
def multiply(a, b, name=None): return _op(lambda x, y: x * y, a, b)
def divide(a, b, name=None): return _op(lambda x, y: x / y, a, b)
def maximum(a, b, name=None): return _op(lambda x, y: torch.maximum(x, torch.as_tensor(y, dtype=x.dtype) if not isinstance(y, torch.Tensor) else y.to(x.dtype)), a, b)
def log(t, name=None): return _op(torch.log, t)
def exp(t, name=None): return _op(torch.exp, t)
def log_sigmoid(t, name=None): return _op(lambda x: -torch.nn.functional.softplus(-x), t)
def shape(t, name=None): return _op(lambda x: torch.as_tensor(list(x.shape), dtype=torch.int64), t)
def cast(t, dtype=None, name=None): return _op(lambda x: x.to(_torch_dtype(dtype)), t)


div = divide


def pow(t, p, name=None):  # noqa: A001  (the TF name)
    return _op(lambda x, y: torch.pow(x, y), t, p)


def concat(values, axis, name=None):
    return Tensor(lambda env: torch.cat([_val(v, env) for v in values], dim=int(axis)))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    def f(x, y):
        x = x.transpose(-1, -2) if transpose_a else x
        y = y.transpose(-1, -2) if transpose_b else y
        return x @ y
    return _op(f, a, b)


def tile(t, multiples, name=None):
    return Tensor(lambda env: _val(t, env).repeat(*[_int(m, env) for m in multiples]))


def transpose(t, perm=None, name=None):
    return _op(lambda x: x.permute(*perm) if perm is not None else x.t(), t)


def matrix_set_diag(t, diagonal, name=None):
    return _op(lambda x, d: x - torch.diag_embed(torch.diagonal(x, dim1=-2, dim2=-1)) + torch.diag_embed(d.to(x.dtype)), t, diagonal)


def _sigmoid_cross_entropy_with_logits(_sentinel=None, labels=None, logits=None, name=None):
    return _op(lambda z, x: torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x))), labels, logits)


nn = types.SimpleNamespace(
    l2_loss=lambda t, name=None: _op(lambda x: (x * x).sum() / 2, t),
    embedding_lookup=lambda params, ids, **k: gather(params, ids),
    sigmoid=lambda t, name=None: _op(torch.sigmoid, t),
    relu=lambda t, name=None: _op(torch.relu, t),
    softmax=lambda t, axis=-1, name=None, dim=None: _op(lambda x: torch.softmax(x, dim=axis if dim is None else dim), t),
    sigmoid_cross_entropy_with_logits=_sigmoid_cross_entropy_with_logits,
)
contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
    xavier_initializer=_xavier_initializer,
    l2_regularizer=lambda scale=0.0, **k: (lambda w: None),    # attached to variables, never added to any loss by the reference
))


# ------------------------------------------------------------------------------------------------ optimizers
class _TrainOp(Tensor):
    def __init__(self, opt, loss):
        self.opt, self.loss, self.variables = opt, loss, list(_VARIABLES)   # tf.trainable_variables() at minimize() time
        Tensor.__init__(self, lambda env: None, "train")

    def apply(self, env):
        loss = self.loss.eval_in(env)
        grads = torch.autograd.grad(loss, [v.value for v in self.variables], allow_unused=True, retain_graph=True)
        with torch.no_grad():
            self.opt._apply([(g, v) for g, v in zip(grads, self.variables) if g is not None])


class _Optimizer(object):
    def minimize(self, loss, **k):
        return _TrainOp(self, loss)


class _SGD(_Optimizer):
    def __init__(self, learning_rate, **k):
        self.lr = float(learning_rate)

    def _apply(self, gv):
        for g, v in gv:
            v.value -= self.lr * g


class _Adagrad(_Optimizer):
    def __init__(self, learning_rate, initial_accumulator_value=0.1, **k):
        self.lr, self.init, self.acc = float(learning_rate), float(initial_accumulator_value), {}

    def _apply(self, gv):
        for g, v in gv:
            acc = self.acc.setdefault(id(v), torch.full_like(v.value, self.init))
            acc += g * g
            v.value -= self.lr * g / torch.sqrt(acc)


class _Adam(_Optimizer):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **k):
        self.lr, self.b1, self.b2, self.eps = float(learning_rate), beta1, beta2, epsilon
        self.t, self.m, self.v = 0, {}, {}

    def _apply(self, gv):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        # the moments of every variable the optimizer owns decay every step (zero gradient = decay only): dense ApplyAdam
        for g, v in gv:
            m = self.m.setdefault(id(v), torch.zeros_like(v.value))
            s = self.v.setdefault(id(v), torch.zeros_like(v.value))
            m += (g - m) * (1 - self.b1)
            s += (g * g - s) * (1 - self.b2)
            v.value -= lr_t * m / (torch.sqrt(s) + self.eps)


class _Saver(object):
    def __init__(self, var_list=None, **k):
        self.var_list = var_list

    def save(self, *a, **k):
        return None

    def restore(self, *a, **k):
        return None


train = types.SimpleNamespace(GradientDescentOptimizer=_SGD, AdagradOptimizer=_Adagrad, AdamOptimizer=_Adam, Saver=_Saver,
                              latest_checkpoint=lambda d: None)


# ------------------------------------------------------------------------------------------------ session
class Session(object):
    class graph(object):
        @staticmethod
        def finalize():
            return None

    def __init__(self, *a, **k):
        pass

    @contextlib.contextmanager
    def as_default(self):
        yield self

    def run(self, fetches, feed_dict=None):
        """Forward values are taken BEFORE the train ops of the same call apply their updates (what `sess.run([train, loss])`
        returns in TF: the loss the gradients were computed from)."""
        env = _Env({id(k): v for k, v in (feed_dict or {}).items()})
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        out = []
        for f in fl:
            if isinstance(f, _TrainOp):
                out.append(None)
                continue
            v = f.eval_in(env)
            out.append(v.detach().numpy().copy() if isinstance(v, torch.Tensor) else v)
        for f in fl:
            if isinstance(f, _TrainOp):
                f.apply(env)
        out = [float(o) if isinstance(o, np.ndarray) and o.ndim == 0 else o for o in out]
        return out[0] if single else out


ConfigProto = lambda *a, **k: types.SimpleNamespace(gpu_options=types.SimpleNamespace())   # noqa: E731
compat = types.SimpleNamespace(v1=types.SimpleNamespace(logging=types.SimpleNamespace(set_verbosity=lambda *a: None, ERROR=0)))


# ------------------------------------------------------------------------------------------------ loading the reference on the shim
def load_reference_models(reference_root, names):
    """Import the genuine reference model modules with THIS module standing in for `tensorflow` (and an empty `gensim`).  Returns
    {name: class}.  sys.modules is restored afterwards, so oracle/refimport.py (which stubs TensorFlow with an inert module for
    the host-only functions) is not disturbed, whichever of the two runs first."""
    me = sys.modules[__name__]
    saved = {k: v for k, v in sys.modules.items() if k in ("tensorflow", "gensim", "gensim.models", "gensim.models.word2vec", "utils", "model")
             or k.startswith(("utils.", "model."))}
    for k in saved:
        del sys.modules[k]
    sys.modules["tensorflow"] = me
    for k in ("gensim", "gensim.models", "gensim.models.word2vec"):
        sys.modules[k] = types.ModuleType(k)
    sys.modules["gensim.models"].Word2Vec = None
    sys.modules["gensim"].models = sys.modules["gensim.models"]
    sys.path.insert(0, reference_root)
    try:
        out = {n: getattr(importlib.import_module("model.ranking." + n), n) for n in names}
    finally:
        sys.path.remove(reference_root)
        for k in [k for k in sys.modules if k in ("tensorflow", "gensim", "gensim.models", "gensim.models.word2vec", "utils", "model")
                  or k.startswith(("utils.", "model."))]:
            del sys.modules[k]
        sys.modules.update(saved)
    return out
