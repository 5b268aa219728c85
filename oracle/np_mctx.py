"""TEST INFRASTRUCTURE — batched NumPy restatement of the mctx search driven by muax.MuZero.act.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product
never does.

PARITY UNPINNED: `mctx` (setup.py:13-20 of the reference, unpinned; 0.0.5 at the reference's commit
date) and `jax.random` are not installable here and the reference ships no tests or golden vectors on
this path.  This module restates the published algorithm as the SAME array
program mctx runs — SoA tree `[B, N, A]`, `vmap`-style masked while-loops that run every tree to the
batch-max depth, one batched recurrent_fn per simulation — and was written independently of the
scalar C restatement (oracle/mz_oracle.c) that it is cross-checked against.

Reference call sites followed:
  muax/model.py:222-243 (_plan), :251-263 (_root_inference), :265-282 (_recurrent_inference)
  muax/nn.py:37-44 (min_max_normalize), :59-115 (Representation / Prediction / Dynamic)
  muax/utils.py:70-76 (_inv_scaling), :94-102 (support_to_scalar)
  muax/policy.py:13-47 (policy kwargs and defaults)
  muax/frameworks/acme/jax/diffusion_muzero/policy.py:62-139 (in-tree mirror of an mctx policy body)

Two math back-ends:
  ExactMath  elementary functions from include/mz_math.h through ctypes and the canonical
             sequential-fma dense layer  -> bit-identical to the C restatement and the CUDA kernels;
  LibmMath   NumPy's own exp/log/expm1 and BLAS matmul -> stands in for "what XLA-CPU would compute";
             used to show that results are stable to ulp-level differences (tolerance 1e-5).
"""
import numpy as np

from . import threefry as tf

F32 = np.float32
TINY = np.finfo(np.float32).tiny
F32_MIN = np.finfo(np.float32).min
UNVISITED = -1
NO_PARENT = -1
ROOT = 0


class LibmMath:
    name = "libm"

    def exp(self, x):
        return np.exp(x, dtype=F32)

    def log(self, x):
        with np.errstate(divide="ignore"):
            return np.log(x, dtype=F32)

    def expm1(self, x):
        return np.expm1(x, dtype=F32)

    def dense(self, x, w, b, onehot=None, num_onehot=0):
        if onehot is not None:
            x = np.concatenate([x, np.eye(num_onehot, dtype=F32)[onehot]], axis=1)
        return (x @ w + b).astype(F32)

    def seq_sum(self, x):
        return x.sum(axis=-1, dtype=F32)


class ExactMath:
    name = "exact"

    def __init__(self):
        from . import c_oracle
        self._c = c_oracle

    def exp(self, x):
        return self._c.expf(x)

    def log(self, x):
        return self._c.logf(x)

    def expm1(self, x):
        return self._c.expm1f(x)

    def dense(self, x, w, b, onehot=None, num_onehot=0):
        """acc = 0; acc = fma(x_k, W_kj, acc) for ascending k; + b.  The one-hot block is concatenated as in
        muax/nn.py:105-108 and goes through the same loop (its zero terms are exact no-ops)."""
        if onehot is not None:
            x = np.concatenate([x, np.eye(num_onehot, dtype=F32)[onehot]], axis=1)
        acc = np.zeros((x.shape[0], w.shape[1]), F32)
        for k in range(x.shape[1]):
            acc = self._c.fmaf(x[:, k:k + 1], w[k:k + 1, :], acc)
        return (acc + b).astype(F32)

    def seq_sum(self, x):
        acc = np.zeros(x.shape[:-1], F32)
        for i in range(x.shape[-1]):
            acc = (acc + x[..., i]).astype(F32)
        return acc


# ---------------------------------------------------------------------------------- nets (muax/nn.py)

def softmax(m, x):
    e = m.exp((x - x.max(axis=-1, keepdims=True)).astype(F32))
    return (e / m.seq_sum(e)[..., None]).astype(F32)


def elu(m, x):
    return np.where(x > 0, x, m.expm1(np.where(x > 0, F32(0), x))).astype(F32)


def min_max_normalize(s):
    s_min = s.min(axis=1, keepdims=True)
    s_max = s.max(axis=1, keepdims=True)
    scale = (s_max - s_min).astype(F32)
    scale = np.where(scale < F32(1e-5), scale + F32(1e-5), scale).astype(F32)
    return ((s - s_min) / scale).astype(F32)


def inv_scaling(x, eps=1e-3):
    x = x.astype(F32)
    t = (np.abs(x) + F32(1) + F32(eps)).astype(F32)
    t = (F32(1) + F32(4 * eps) * t).astype(F32)
    t = ((np.sqrt(t, dtype=F32) - F32(1)) / F32(2 * eps)).astype(F32)
    return (np.sign(x) * ((t * t).astype(F32) - F32(1))).astype(F32)


def support_to_scalar(m, probs, support_size):
    rng = (np.arange(2 * support_size + 1) - support_size).astype(F32)
    return inv_scaling(m.seq_sum((rng * probs).astype(F32)))


def stack(m, layers, x, act, onehot=None, num_onehot=0):
    for i, (w, b) in enumerate(layers):
        x = m.dense(x, w, b, onehot if i == 0 else None, num_onehot)
        if i != len(layers) - 1:
            x = elu(m, x) if act == 0 else np.maximum(x, F32(0))
    return x


class Model:
    """muax.MuZero's inference half over declarative MLP stacks."""

    def __init__(self, nets, math, support_size=10, discount=0.99, activation=0, repr_minmax=True, dyn_minmax=True):
        self.nets = {k: [(np.asarray(w, F32), np.asarray(b, F32)) for w, b in v] for k, v in nets.items()}
        self.m = math
        self.S = support_size
        self.discount = F32(discount)
        self.act = activation
        self.repr_minmax = repr_minmax
        self.dyn_minmax = dyn_minmax
        self.num_actions = self.nets["pred_pi"][-1][0].shape[1]

    def root_inference(self, obs):  # muax/model.py:251-263
        s = stack(self.m, self.nets["repr"], obs.astype(F32), self.act)
        if self.repr_minmax:
            s = min_max_normalize(s)
        v = stack(self.m, self.nets["pred_v"], s, self.act)
        logits = stack(self.m, self.nets["pred_pi"], s, self.act)
        v = support_to_scalar(self.m, softmax(self.m, v), self.S)
        return logits, v, s

    def recurrent_inference(self, action, emb):  # muax/model.py:265-282
        A = self.num_actions
        r = stack(self.m, self.nets["dyn_r"], emb, self.act, action, A)
        ns = stack(self.m, self.nets["dyn_ns"], emb, self.act, action, A)
        if self.dyn_minmax:
            ns = min_max_normalize(ns)
        v = stack(self.m, self.nets["pred_v"], ns, self.act)
        logits = stack(self.m, self.nets["pred_pi"], ns, self.act)
        r = support_to_scalar(self.m, softmax(self.m, r), self.S)
        v = support_to_scalar(self.m, softmax(self.m, v), self.S)
        discount = np.ones_like(r) * self.discount
        return r, discount, logits, v, ns


# ---------------------------------------------------------------------------------- tree (A.1)

class Tree:
    def __init__(self, B, N, A, E):
        self.node_visits = np.zeros((B, N), np.int32)
        self.raw_values = np.zeros((B, N), F32)
        self.node_values = np.zeros((B, N), F32)
        self.parents = np.full((B, N), NO_PARENT, np.int32)
        self.action_from_parent = np.full((B, N), NO_PARENT, np.int32)
        self.children_index = np.full((B, N, A), UNVISITED, np.int32)
        self.children_prior_logits = np.zeros((B, N, A), F32)
        self.children_values = np.zeros((B, N, A), F32)
        self.children_visits = np.zeros((B, N, A), np.int32)
        self.children_rewards = np.zeros((B, N, A), F32)
        self.children_discounts = np.zeros((B, N, A), F32)
        self.embeddings = np.zeros((B, N, E), F32)
        self.root_invalid_actions = None
        self.root_gumbel = None

    def qvalues(self, rows, node):
        return (self.children_rewards[rows, node]
                + (self.children_discounts[rows, node] * self.children_values[rows, node]).astype(F32)).astype(F32)


def update_tree_node(tree, rows, node, prior_logits, value, emb):
    tree.node_visits[rows, node] += 1
    tree.children_prior_logits[rows, node] = prior_logits
    tree.raw_values[rows, node] = value
    tree.node_values[rows, node] = value
    tree.embeddings[rows, node] = emb


# ---------------------------------------------------------------------------------- qtransforms (A.6)

def qtransform_by_parent_and_siblings(m, tree, rows, node, epsilon=1e-8):
    q = tree.qvalues(rows, node)
    vc = tree.children_visits[rows, node]
    nv = tree.node_values[rows, node][:, None]
    safe = np.where(vc > 0, q, nv)
    lo = np.minimum(nv, safe.min(axis=-1, keepdims=True))
    hi = np.maximum(nv, safe.max(axis=-1, keepdims=True))
    completed = np.where(vc > 0, q, lo)
    return ((completed - lo) / np.maximum((hi - lo).astype(F32), F32(epsilon))).astype(F32)


def qtransform_completed_by_mix_value(m, tree, rows, node, value_scale=0.1, maxvisit_init=50.0, epsilon=1e-8):
    q = tree.qvalues(rows, node)
    vc = tree.children_visits[rows, node]
    raw = tree.raw_values[rows, node]
    p = np.maximum(TINY, softmax(m, tree.children_prior_logits[rows, node]))
    sum_vc = vc.sum(axis=-1)
    sum_p = m.seq_sum(np.where(vc > 0, p, F32(0)))
    terms = np.where(vc > 0, (p * q).astype(F32) / np.where(vc > 0, sum_p[:, None], F32(1)), F32(0)).astype(F32)
    weighted_q = m.seq_sum(terms)
    mixed = ((raw + (sum_vc.astype(F32) * weighted_q).astype(F32)).astype(F32) / (sum_vc + 1).astype(F32)).astype(F32)
    completed = np.where(vc > 0, q, mixed[:, None]).astype(F32)
    lo = completed.min(axis=-1, keepdims=True)
    hi = completed.max(axis=-1, keepdims=True)
    completed = ((completed - lo) / np.maximum((hi - lo).astype(F32), F32(epsilon))).astype(F32)
    visit_scale = (F32(maxvisit_init) + vc.max(axis=-1).astype(F32)).astype(F32)
    return ((visit_scale * F32(value_scale)).astype(F32)[:, None] * completed).astype(F32)


QTRANSFORMS = {0: qtransform_by_parent_and_siblings, 1: qtransform_completed_by_mix_value}


def masked_argmax(x, invalid):
    if invalid is not None:
        x = np.where(invalid.astype(bool), -np.inf, x)
    return np.argmax(x, axis=-1).astype(np.int32)


# ---------------------------------------------------------------------------------- action selection (A.4, A.5)

def muzero_policy_score(m, tree, rows, node, pb_c_init, pb_c_base):
    """The exploration term of pUCT (A.5): sqrt(n) * pb_c(n) * prior / (visits + 1).  In-tree cross-check of the
    formula: muax/frameworks/acme/tf/mcts/search.py:463-497 `puct` (tests/test_oracle_golden.py::test_puct_pins)."""
    vc = tree.children_visits[rows, node]
    node_visit = tree.node_visits[rows, node].astype(F32)
    pb_c = (F32(pb_c_init) + m.log((((node_visit + F32(pb_c_base)).astype(F32) + F32(1)) / F32(pb_c_base)).astype(F32)))
    probs = softmax(m, tree.children_prior_logits[rows, node])
    return ((((np.sqrt(node_visit, dtype=F32) * pb_c).astype(F32)[:, None] * probs).astype(F32))
            / (vc + 1).astype(F32)).astype(F32)


def muzero_action_selection(m, keys, tree, rows, node, depth, qt, pb_c_init, pb_c_base, mode):
    A = tree.children_visits.shape[-1]
    policy_score = muzero_policy_score(m, tree, rows, node, pb_c_init, pb_c_base)
    value_score = qt(m, tree, rows, node)
    noise = (F32(1e-7) * tf.uniform(keys, A, mode)).astype(F32)
    to_argmax = ((value_score + policy_score).astype(F32) + noise).astype(F32)
    invalid = tree.root_invalid_actions[rows] * (depth[:, None] == 0)
    return masked_argmax(to_argmax, invalid)


def considered_visits_sequence(max_considered, num_simulations):
    if max_considered <= 1:
        return tuple(range(num_simulations))
    log2max = int(np.ceil(np.log2(max_considered)))
    seq, visits, k = [], [0] * max_considered, max_considered
    while len(seq) < num_simulations:
        extra = max(1, int(num_simulations / (log2max * k)))
        for _ in range(extra):
            seq.extend(visits[:k])
            for i in range(k):
                visits[i] += 1
        k = max(2, k // 2)
    return tuple(seq[:num_simulations])


def considered_visits_table(max_considered, num_simulations):
    return np.array([considered_visits_sequence(mm, num_simulations) for mm in range(max_considered + 1)],
                    np.int32).reshape(max_considered + 1, num_simulations)


def score_considered(considered_visit, gumbel, logits, q, visits):
    logits = (logits - logits.max(axis=-1, keepdims=True)).astype(F32)
    penalty = np.where(visits == considered_visit, F32(0), -np.inf).astype(F32)
    return (np.maximum(F32(-1e9), ((gumbel + logits).astype(F32) + q).astype(F32)) + penalty).astype(F32)


def gumbel_root_action_selection(m, tree, rows, qt, table, max_considered):
    vc = tree.children_visits[rows, ROOT]
    logits = tree.children_prior_logits[rows, ROOT]
    q = qt(m, tree, rows, np.zeros_like(rows))
    num_valid = (1 - tree.root_invalid_actions[rows].astype(np.int32)).sum(axis=-1)
    num_considered = np.minimum(max_considered, num_valid)
    sim_index = vc.sum(axis=-1)
    considered_visit = table[num_considered, sim_index]
    to_argmax = score_considered(considered_visit[:, None], tree.root_gumbel[rows], logits, q, vc)
    return masked_argmax(to_argmax, tree.root_invalid_actions[rows])


def gumbel_interior_action_selection(m, tree, rows, node, qt):
    vc = tree.children_visits[rows, node]
    logits = tree.children_prior_logits[rows, node]
    q = qt(m, tree, rows, node)
    probs = softmax(m, (logits + q).astype(F32))
    to_argmax = (probs - (vc.astype(F32) / (1 + vc.sum(axis=-1, keepdims=True)).astype(F32)).astype(F32)).astype(F32)
    return np.argmax(to_argmax, axis=-1).astype(np.int32)


# ---------------------------------------------------------------------------------- search (A.3)

def simulate(m, keys, tree, select_fn, max_depth, mode):
    """vmap(while_loop): every tree steps until the slowest one in the batch is done; finished rows are frozen."""
    B = tree.node_visits.shape[0]
    keys = keys.copy()
    node_index = np.full(B, NO_PARENT, np.int32)
    action = np.full(B, NO_PARENT, np.int32)
    next_node = np.zeros(B, np.int32)
    depth = np.zeros(B, np.int32)
    cont = np.ones(B, bool)
    while cont.any():
        rows = np.nonzero(cont)[0]
        sp = tf.split(keys[rows], 2, mode)
        keys[rows] = sp[:, 0]
        node = next_node[rows]
        a = select_fn(sp[:, 1], rows, node, depth[rows])
        nxt = tree.children_index[rows, node, a]
        node_index[rows] = node
        action[rows] = a
        next_node[rows] = nxt
        depth[rows] += 1
        cont[rows] = (nxt != UNVISITED) & (depth[rows] < max_depth)
    return node_index, action, depth


def expand(model, tree, parent, action, nxt):
    B = parent.shape[0]
    rows = np.arange(B)
    emb = tree.embeddings[rows, parent]
    r, discount, logits, v, new_emb = model.recurrent_inference(action, emb)
    update_tree_node(tree, rows, nxt, logits, v, new_emb)
    tree.children_index[rows, parent, action] = nxt
    tree.children_rewards[rows, parent, action] = r
    tree.children_discounts[rows, parent, action] = discount
    tree.parents[rows, nxt] = parent
    tree.action_from_parent[rows, nxt] = action


def backward(tree, leaf):
    B = leaf.shape[0]
    index = leaf.astype(np.int32).copy()
    G = tree.node_values[np.arange(B), index].copy()
    while True:
        rows = np.nonzero(index != ROOT)[0]
        if rows.size == 0:
            break
        idx = index[rows]
        p = tree.parents[rows, idx]
        count = tree.node_visits[rows, p]
        a = tree.action_from_parent[rows, idx]
        reward = tree.children_rewards[rows, p, a]
        g = (reward + (tree.children_discounts[rows, p, a] * G[rows]).astype(F32)).astype(F32)
        G[rows] = g
        cf = count.astype(F32)
        parent_value = (((tree.node_values[rows, p] * cf).astype(F32) + g).astype(F32) / (cf + F32(1))).astype(F32)
        child_value = tree.node_values[rows, idx]
        tree.node_values[rows, p] = parent_value
        tree.node_visits[rows, p] = count + 1
        tree.children_values[rows, p, a] = child_value
        tree.children_visits[rows, p, a] += 1
        index[rows] = p


def search(model, key, root, num_simulations, max_depth, invalid_actions, root_sel, interior_sel, mode,
           root_gumbel=None, global_batch=None, batch_offset=0):
    logits, value, emb = root
    B, A = logits.shape
    GB = B if global_batch is None else global_batch
    if max_depth is None:
        max_depth = num_simulations
    tree = Tree(B, num_simulations + 1, A, emb.shape[1])
    tree.root_invalid_actions = (np.zeros((B, A), np.uint8) if invalid_actions is None
                                 else np.asarray(invalid_actions).astype(np.uint8))
    tree.root_gumbel = root_gumbel
    rows_all = np.arange(B)
    update_tree_node(tree, rows_all, np.zeros(B, np.int32), logits, value, emb)
    depths = np.zeros((B, num_simulations), np.int32)

    def select_fn(keys, rows, node, depth):
        is_root = depth == 0
        a = np.zeros(rows.shape[0], np.int32)
        if is_root.any():
            a[is_root] = root_sel(keys[is_root], tree, rows[is_root], node[is_root], depth[is_root])
        if (~is_root).any():
            ni = ~is_root
            a[ni] = interior_sel(keys[ni], tree, rows[ni], node[ni], depth[ni])
        return a

    rng = np.asarray(key, np.uint32)
    for sim in range(num_simulations):
        rng, simulate_key, _expand_key = tf.split(rng, 3, mode)
        simulate_keys = tf.split(simulate_key, GB, mode)[batch_offset:batch_offset + B]
        parent, action, depth = simulate(model.m, simulate_keys, tree, select_fn, max_depth, mode)
        depths[:, sim] = depth
        nxt = tree.children_index[rows_all, parent, action]
        nxt = np.where(nxt == UNVISITED, sim + 1, nxt).astype(np.int32)
        expand(model, tree, parent, action, nxt)
        backward(tree, nxt)
    return tree, depths


# ---------------------------------------------------------------------------------- policies (A.2, A.4)

def mask_invalid_actions(logits, invalid):
    if invalid is None:
        return logits
    logits = (logits - logits.max(axis=-1, keepdims=True)).astype(F32)
    return np.where(np.asarray(invalid).astype(bool), F32_MIN, logits).astype(F32)


def summary_visit_probs(tree):
    A = tree.children_visits.shape[-1]
    vc = tree.children_visits[:, ROOT].astype(F32)
    total = np.zeros(vc.shape[0], F32)
    for a in range(A):
        total = (total + vc[:, a]).astype(F32)
    total = total[:, None]
    probs = (vc / np.maximum(total, F32(1))).astype(F32)
    return np.where(total > 0, probs, F32(1 / A)).astype(F32)


def muzero_policy(model, key, root, num_simulations, invalid_actions=None, max_depth=None, qtransform=0,
                  dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652.0, temperature=1.0,
                  mode=tf.LEGACY, dirichlet_noise=None, global_batch=None, batch_offset=0):
    m = model.m
    logits, value, emb = root
    B, A = logits.shape
    GB = B if global_batch is None else global_batch
    rng_key, _dirichlet_key, search_key = tf.split(np.asarray(key, np.uint32), 3, mode)
    probs = softmax(m, logits)
    if dirichlet_noise is None:  # framework-defined sampler (jax.random.dirichlet is not reproducible off-XLA)
        from . import c_oracle
        dirichlet_noise = c_oracle.dirichlet(_dirichlet_key, batch_offset, B, A, dirichlet_alpha)
    noise = np.asarray(dirichlet_noise, F32)
    noisy = ((F32(1) - F32(dirichlet_fraction)) * probs + F32(dirichlet_fraction) * noise).astype(F32)
    new_logits = m.log(np.maximum(noisy, TINY))
    new_logits = mask_invalid_actions(new_logits, invalid_actions)
    qt = QTRANSFORMS[qtransform]

    def sel(keys, tree, rows, node, depth):
        return muzero_action_selection(m, keys, tree, rows, node, depth, qt, pb_c_init, pb_c_base, mode)

    tree, depths = search(model, search_key, (new_logits, value, emb), num_simulations, max_depth, invalid_actions,
                          sel, sel, mode, global_batch=GB, batch_offset=batch_offset)
    w = summary_visit_probs(tree)
    l = m.log(np.maximum(w, TINY))
    l = (l - l.max(axis=-1, keepdims=True)).astype(F32)
    l = (l / np.maximum(TINY, F32(temperature))).astype(F32)
    u = tf.uniform_tiny(rng_key, GB * A, mode).reshape(GB, A)[batch_offset:batch_offset + B]
    g = (-m.log((-m.log(u)).astype(F32))).astype(F32)
    action = np.argmax((g + l).astype(F32), axis=-1).astype(np.int32)
    return dict(action=action, action_weights=w, tree=tree, sim_depth=depths, root_noise=noise)


def gumbel_muzero_policy(model, key, root, num_simulations, invalid_actions=None, max_depth=None, qtransform=1,
                         max_num_considered_actions=16, gumbel_scale=1.0, mode=tf.LEGACY, root_gumbel=None,
                         global_batch=None, batch_offset=0):
    m = model.m
    logits, value, emb = root
    B, A = logits.shape
    GB = B if global_batch is None else global_batch
    logits = mask_invalid_actions(logits, invalid_actions)
    rng_key, gumbel_key = tf.split(np.asarray(key, np.uint32), 2, mode)
    if root_gumbel is None:
        u = tf.uniform_tiny(gumbel_key, GB * A, mode).reshape(GB, A)[batch_offset:batch_offset + B]
        root_gumbel = (F32(gumbel_scale) * (-m.log((-m.log(u)).astype(F32))).astype(F32)).astype(F32)
    gumbel = np.asarray(root_gumbel, F32)
    qt = QTRANSFORMS[qtransform]
    table = considered_visits_table(max_num_considered_actions, num_simulations)

    def root_sel(keys, tree, rows, node, depth):
        return gumbel_root_action_selection(m, tree, rows, qt, table, max_num_considered_actions)

    def interior_sel(keys, tree, rows, node, depth):
        return gumbel_interior_action_selection(m, tree, rows, node, qt)

    tree, depths = search(model, rng_key, (logits, value, emb), num_simulations, max_depth, invalid_actions, root_sel,
                          interior_sel, mode, root_gumbel=gumbel, global_batch=GB, batch_offset=batch_offset)
    rows = np.arange(B)
    vc = tree.children_visits[:, ROOT]
    considered_visit = vc.max(axis=-1, keepdims=True)
    completed_q = qt(m, tree, rows, np.zeros(B, np.int32))
    to_argmax = score_considered(considered_visit, gumbel, logits, completed_q, vc)
    action = masked_argmax(to_argmax, invalid_actions)
    w = softmax(m, mask_invalid_actions((logits + completed_q).astype(F32), invalid_actions))
    return dict(action=action, action_weights=w, tree=tree, sim_depth=depths, root_noise=gumbel)


# ---------------------------------------------------------------------------------- stochastic MuZero
# mctx.stochastic_muzero_policy (call site: muax/policy.py:50-67; skeleton of the afterstate search visible in-tree
# at muax/frameworks/acme/jax/diffusion_muzero/policy.py:77-129,150-211).  PARITY UNPINNED like the rest of this
# module (mctx is absent): restated from the published algorithm — the tree holds A' = A + C pseudo-actions; the
# embedding of a node is (state, afterstate, is_decision); `stochastic_recurrent_fn` evaluates BOTH the decision and the
# chance function for every row and keeps one by the PARENT's node type; decision nodes select with pUCT over all A'
# slots (chance slots carry -inf priors), chance nodes with argmax softmax(prior) / (visits + 1); Dirichlet noise,
# the invalid-action mask, the visit summary and the final draw see the A decision actions only.  Root invalid actions
# are padded with zeros for the chance slots (their -inf prior already keeps them out).

class StochasticModel:
    """Adapter with Model's `recurrent_inference(action, emb)` surface over a (decision_fn, chance_fn) pair.
    decision_fn(action[B], state[B, Es]) -> (chance_logits[B, C], afterstate_value[B], afterstate[B, Ea]);
    chance_fn(outcome[B], afterstate[B, Ea]) -> (action_logits[B, A], value[B], reward[B], discount[B], state[B, Es]).
    Flat embedding layout: [state (Es) | afterstate (Ea) | is_decision (1)]."""

    def __init__(self, math, decision_fn, chance_fn, num_actions, num_chance, state_dim, afterstate_dim):
        self.m = math
        self.decision_fn, self.chance_fn = decision_fn, chance_fn
        self.A, self.C, self.Es, self.Ea = num_actions, num_chance, state_dim, afterstate_dim

    def root_embedding(self, state):
        B = state.shape[0]
        _, _, after = self.decision_fn(np.zeros(B, np.int32), state)  # mctx builds a dummy afterstate the same way
        return np.concatenate([state, after, np.ones((B, 1), F32)], axis=1).astype(F32)

    def recurrent_inference(self, action, emb):
        B = emb.shape[0]
        state, after, is_dec = emb[:, :self.Es], emb[:, self.Es:self.Es + self.Ea], emb[:, -1] != 0
        # mctx hands both functions the raw pseudo-action and discards one result; the index a row does not use is
        # clamped into range here (XLA's gather clamps out-of-range indices silently)
        chance_logits, after_value, new_after = self.decision_fn(np.minimum(action, self.A - 1), state)
        act_logits, value, reward, discount, new_state = self.chance_fn(np.maximum(action - self.A, 0), after)
        ninf = lambda n: np.full((B, n), -np.inf, F32)  # noqa: E731
        logits = np.where(is_dec[:, None], np.concatenate([ninf(self.A), chance_logits], 1),
                          np.concatenate([act_logits, ninf(self.C)], 1)).astype(F32)
        v = np.where(is_dec, after_value, value).astype(F32)
        r = np.where(is_dec, F32(0), reward).astype(F32)
        d = np.where(is_dec, F32(1), discount).astype(F32)
        new_emb = np.concatenate([new_state, new_after, (~is_dec).astype(F32)[:, None]], axis=1).astype(F32)
        return r, d, logits, v, new_emb


def stochastic_muzero_policy(model, key, root, num_simulations, invalid_actions=None, max_depth=None, qtransform=0,
                             dirichlet_fraction=0.25, dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652.0,
                             temperature=1.0, mode=tf.LEGACY, dirichlet_noise=None):
    """root = (prior_logits[B, A], value[B], state[B, Es]); model = StochasticModel."""
    m = model.m
    logits, value, state = root
    B, A = logits.shape
    C = model.C
    rng_key, _dirichlet_key, search_key = tf.split(np.asarray(key, np.uint32), 3, mode)
    probs = softmax(m, logits)
    if dirichlet_noise is None:
        from . import c_oracle
        dirichlet_noise = c_oracle.dirichlet(_dirichlet_key, 0, B, A, dirichlet_alpha)
    noise = np.asarray(dirichlet_noise, F32)
    noisy = ((F32(1) - F32(dirichlet_fraction)) * probs + F32(dirichlet_fraction) * noise).astype(F32)
    new_logits = mask_invalid_actions(m.log(np.maximum(noisy, TINY)), invalid_actions)
    new_logits = np.concatenate([new_logits, np.full((B, C), -np.inf, F32)], axis=1)
    invalid_padded = None if invalid_actions is None else np.concatenate(
        [np.asarray(invalid_actions).astype(np.uint8), np.zeros((B, C), np.uint8)], axis=1)
    qt = QTRANSFORMS[qtransform]

    def sel(keys, tree, rows, node, depth):
        decision = muzero_action_selection(m, keys, tree, rows, node, depth, qt, pb_c_init, pb_c_base, mode)
        prob = softmax(m, tree.children_prior_logits[rows, node])
        chance = np.argmax((prob / (tree.children_visits[rows, node] + 1).astype(F32)).astype(F32), axis=-1)
        is_dec = tree.embeddings[rows, node, -1] != 0
        return np.where(is_dec, decision, chance).astype(np.int32)

    with np.errstate(invalid="ignore"):  # -inf - -inf inside softmax of fully masked slots never happens; inf * 0 may
        tree, depths = search(model, search_key, (new_logits, value, model.root_embedding(np.asarray(state, F32))),
                              num_simulations, max_depth, invalid_padded, sel, sel, mode)
    vc = tree.children_visits[:, ROOT, :A].astype(F32)
    total = np.zeros(B, F32)
    for a in range(A):
        total = (total + vc[:, a]).astype(F32)
    total = total[:, None]
    w = np.where(total > 0, (vc / np.maximum(total, F32(1))).astype(F32), F32(1 / A)).astype(F32)
    l = m.log(np.maximum(w, TINY))
    l = (l - l.max(axis=-1, keepdims=True)).astype(F32)
    l = (l / np.maximum(TINY, F32(temperature))).astype(F32)
    u = tf.uniform_tiny(rng_key, B * A, mode).reshape(B, A)
    g = (-m.log((-m.log(u)).astype(F32))).astype(F32)
    action = np.argmax((g + l).astype(F32), axis=-1).astype(np.int32)
    return dict(action=action, action_weights=w, tree=tree, sim_depth=depths, root_noise=noise)


def act(nets, key, obs=None, root=None, math=None, policy=0, invalid=None, noise=None, support_size=10,
        discount=0.99, activation=0, repr_minmax=1, dyn_minmax=1, prng_mode=tf.LEGACY, num_simulations=5,
        max_depth=None, qtransform=0, max_considered=16, gumbel_scale=1.0, temperature=1.0, dirichlet_fraction=0.25,
        dirichlet_alpha=0.3, pb_c_init=1.25, pb_c_base=19652.0, global_batch=None, batch_offset=0):
    """muax.MuZero._plan (model.py:222-243) with the same keyword surface as oracle.c_oracle.search."""
    math = math or ExactMath()
    model = Model(nets, math, support_size, discount, activation, bool(repr_minmax), bool(dyn_minmax))
    if obs is not None:
        root = model.root_inference(np.asarray(obs, F32))
    else:
        root = tuple(np.asarray(x, F32) for x in root)
    max_depth = None if not max_depth else max_depth
    if policy == 0:
        out = muzero_policy(model, key, root, num_simulations, invalid, max_depth, qtransform, dirichlet_fraction,
                            dirichlet_alpha, pb_c_init, pb_c_base, temperature, prng_mode, noise, global_batch,
                            batch_offset)
    else:
        out = gumbel_muzero_policy(model, key, root, num_simulations, invalid, max_depth, qtransform, max_considered,
                                   gumbel_scale, prng_mode, noise, global_batch, batch_offset)
    tree = out.pop("tree")
    for name in ("node_visits", "parents", "action_from_parent", "children_index", "children_visits", "raw_values",
                 "node_values", "children_prior_logits", "children_values", "children_rewards", "children_discounts",
                 "embeddings"):
        out[name] = getattr(tree, name)
    out["root_value"] = root[1]  # raw network value, muax/model.py:243
    return out
