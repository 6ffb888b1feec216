"""CPU ORACLE (numpy, fp32) for the EVRP pointer-policy rollout hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it, and only as the checker
or as the timed CPU baseline.  The product path (the ``evrp_b200`` package) never imports it and
fails loudly when its CUDA library is missing.

It restates, function by function, the reference's algorithm (citations are relative to
/root/reference/EM-EVRP unless they say DM-EVRP):

  pairwise_matrices      problems/EVRP.py:85-92 and nets/DRLModel.py:124-133
  init_embed             nets/DRLModel.py:192-200
  encoder                nets/GraphEncoder.py:55-118 (MHA), :142-150 (BatchNorm1d), :153-179, :202-214
  precompute             nets/DRLModel.py:345-362, :458-465
  update_dynamic         problems/EVRP.py:226-280   (DM variant: DM-EVRP/problems/EVRP.py:231,280)
  update_mask            problems/EVRP.py:105-223
  route_context          nets/DRLModel.py:203-237   (running max instead of the O(T^2) re-gather;
                                                    max is exact and associative => bit-identical)
  step_log_p             nets/DRLModel.py:379-452
  rollout                nets/DRLModel.py:240-280, :316-342, :182-190, :109-140
  sample_many            nets/DRLModel.py:283-313
  reinforce_loss         trainer.py:115-122

Parity pin: the reference ships no tests / golden vectors, so this restatement
is pinned against OUTPUTS OF THE PATCHED REFERENCE ITSELF, generated in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (state/reward/mask bit-exact,
log_p to 2e-5, tours equal except audited near-ties) — see tests/test_oracle_golden.py.

Arithmetic convention: every binary operation of the state/mask recipe is one separately rounded
fp32 operation in the reference's order; ``x / python_scalar`` is a true IEEE division (the CPU
convention of ATen).
"""
import math

import numpy as np

f32 = np.float32


# ----------------------------------------------------------------------------------------------
# constants of the vehicle / problem (problems/EVRP.py:29-41, trainer.py:225-228)
# ----------------------------------------------------------------------------------------------
class EVRPConsts:
    def __init__(self, charging_num=5, velocity=50., Start_SOC=80., t_limit=10., max_load=4):
        self.F = int(charging_num)
        self.velocity = float(velocity)
        self.Start_SOC = float(Start_SOC)
        self.t_limit = float(t_limit)
        self.max_load = max_load
        self.mc, self.g, self.w = 4100, 9.81, 1000
        self.Cd, self.A, self.Ad, self.Cr = 0.7, 6.66, 1.2041, 0.01
        self.motor_d, self.motor_r, self.battery_d, self.battery_r = 1.18, 0.85, 1.11, 0.93
        self.charging_time, self.serve_time = 1, 0.33
        # python-double sub-expressions that the reference folds before touching a tensor
        self.aero = f32(0.5 * self.Cd * self.A * self.Ad * (self.velocity / 3.6) ** 2)
        self.kd = f32(self.motor_d * self.battery_d)
        self.kr = f32(self.motor_r * self.battery_r)


def _energy(c, mass, slope, tc):
    """soc consumption of one leg; problems/EVRP.py:247-252 (and the 4 copies in update_mask)."""
    mg = mass * f32(c.g)
    Pm = ((c.aero + mg * slope) + mg * f32(c.Cr)) * f32(c.velocity)
    pos = ((c.kd * Pm) * tc) / f32(3600.)
    neg = ((c.kr * Pm) * tc) / f32(3600.)
    out = np.zeros_like(Pm)
    out = np.where(Pm > 0, pos, out)
    out = np.where(Pm < 0, neg, out)
    return out.astype(f32)


# ----------------------------------------------------------------------------------------------
# instance side
# ----------------------------------------------------------------------------------------------
def pairwise_matrices(static, elevation):
    """static [B,2,S] (coords), elevation [B,S] -> distances [B,S,S], slope [B,S,S] (fp32)."""
    static = static.astype(f32)
    elevation = elevation.astype(f32)
    diff = static[:, :, :, None] - static[:, :, None, :]          # [B,2,S,S]  (i, j)
    sq = diff * diff
    D = np.sqrt(sq[:, 0] + sq[:, 1]).astype(f32)
    dE = elevation[:, :, None] - elevation[:, None, :]
    G = np.clip(dE / (D + f32(0.000001)), f32(-0.10), f32(0.10)).astype(f32)
    return D, G


def generate_instances(B, N, F=5, seed=0, t_limit=10., Start_SOC=80., max_load=4, max_demand=4):
    """Synthetic instances with the distribution of problems/EVRP.py:43-92 (layout after patch P1:
    [depot, F stations, N customers]).  Uses numpy's generator, so the *stream* differs from the
    reference's torch RNG; the distribution and value grid (integers, demand in {1..4}/16,
    elevation in {0..100}/1000) are the reference's."""
    rng = np.random.RandomState(seed)
    S = N + F + 1
    x1 = rng.randint(0, 50, (B, F)); x2 = rng.randint(51, 101, (B, F))
    y1 = rng.randint(0, 50, (B, F)); y2 = rng.randint(51, 101, (B, F))
    st = np.zeros((B, 2, F), f32)
    for i in range(F):
        st[:, 0, i] = (x1 if i % 4 in (0, 1) else x2)[:, i]
        st[:, 1, i] = (y1 if i % 4 in (0, 2) else y2)[:, i]
    depot = rng.randint(25, 75, (B, 2, 1)).astype(f32)
    cust = rng.randint(0, 101, (B, 2, N)).astype(f32)
    static = np.concatenate([depot, st, cust], 2).astype(f32)
    demand = (rng.randint(1, max_demand + 1, (B, S)).astype(f32) * f32(0.25)) / f32(max_load)
    demand[:, :F + 1] = 0
    elevation = (rng.randint(0, 101, (B, S)).astype(f32) / f32(1000))
    dynamic = np.stack([np.full((B, S), 1., f32), demand.astype(f32),
                        np.full((B, S), Start_SOC, f32), np.full((B, S), t_limit, f32)], 1)
    return static, dynamic, elevation


# ----------------------------------------------------------------------------------------------
# environment: update_dynamic / update_mask
# ----------------------------------------------------------------------------------------------
def update_dynamic(c, dynamic, D, G, now, chosen):
    """-> new_dynamic [B,4,S], energy [B,1], distance [B,1].  problems/EVRP.py:226-280."""
    B = dynamic.shape[0]
    ar = np.arange(B)
    F = c.F
    dist = D[ar, now, chosen][:, None]
    sl = G[ar, now, chosen][:, None]
    depot = chosen == 0
    station = (chosen > 0) & (chosen <= F)
    station_depot = chosen <= F
    customer = chosen > F
    loads = dynamic[:, 0].copy(); dem = dynamic[:, 1].copy()
    soc = dynamic[:, 2].copy(); tim = dynamic[:, 3].copy()
    tc = dist / f32(c.velocity)
    mass = (f32(c.mc) + (loads[:, 0] * f32(c.max_load)) * f32(c.w))[:, None]
    e = _energy(c, mass, sl, tc)
    tim -= tc
    tim[depot] = f32(c.t_limit)
    tim[station] -= f32(c.charging_time)
    tim[customer] -= f32(c.serve_time)
    soc -= e
    soc[station_depot] = f32(c.Start_SOC)
    load = loads[ar, chosen]
    dm = dem[ar, chosen]
    if customer.any():
        new_load = np.maximum(load - dm, f32(0))
        new_dem = np.maximum(dm - load, f32(0))
        ci = np.nonzero(customer)[0]
        loads[ci] = new_load[ci][:, None]
        dem[ci, chosen[ci]] = new_dem[ci]
        dem[ci, 0] = f32(-1.) + new_load[ci]
    if depot.any():
        di = np.nonzero(depot)[0]
        loads[di] = f32(1.)
        dem[di, 0] = f32(0.)
    new_dyn = np.stack([loads, dem, soc, tim], 1).astype(f32)
    return new_dyn, e.astype(f32), dist.astype(f32)


def station_constants(c, D, G):
    """Instance constants of masking rule 7 (problems/EVRP.py:196-202): nearest station of every
    node (first index on ties), its distance d3 (d3[0] forced to 0), the slope station->node s3 and
    the scalar d4 = D[nearest_station(depot), depot]."""
    F = c.F
    Ds = D[:, 1:F + 1, :]
    k = np.argmin(Ds, 1)                                   # [B,S]
    d3 = np.take_along_axis(Ds, k[:, None, :], 1)[:, 0].copy()
    d3[:, 0] = 0
    s3 = np.take_along_axis(G[:, 1:F + 1, :], k[:, None, :], 1)[:, 0]
    d4 = np.take_along_axis(D[:, 1:F + 1, 0], k[:, 0:1], 1)   # [B,1]
    return k, d3.astype(f32), s3.astype(f32), d4.astype(f32)


def update_mask(c, dynamic, D, G, chosen, global_done=None):
    """-> mask [B,S] fp32 0/1.  problems/EVRP.py:105-223.

    ``global_done``: the reference returns an all-zero mask only when EVERY instance of the batch
    has zero demand everywhere (:127-128).  None = evaluate that test on this batch."""
    B, _, S = dynamic.shape
    F = c.F
    ar = np.arange(B)
    loads, dem, soc, tim = dynamic[:, 0], dynamic[:, 1], dynamic[:, 2], dynamic[:, 3]
    if global_done is None:
        global_done = bool((dem == 0).all())
    if global_done:
        return np.zeros((B, S), f32)
    depot = chosen == 0
    station = (chosen > 0) & (chosen <= F)
    station_depot = chosen <= F
    customer = chosen > F
    m = (dem != 0) & (dem < loads)                                    # rule 1
    m[ar, chosen] = False                                             # rule 2
    m[station_depot, 1:F + 1] = False                                 # rule 3
    m[station, 0] = True
    m[customer, :F + 1] = True
    m[depot, 0] = False
    combined = (loads[:, 0] == 0) | (dem[:, F:].sum(1) == 0)          # rule 4
    m[combined, 0] = True
    m[combined, 1:] = False
    d0 = D[ar, chosen]                                                # rule 5
    g0 = G[ar, chosen]
    tc0 = d0 / f32(c.velocity)
    mass0 = (f32(c.mc) + (loads[:, 0] * f32(c.max_load)) * f32(c.w))[:, None]
    soc0 = _energy(c, np.broadcast_to(mass0, d0.shape), g0, tc0)
    mass2 = f32(c.mc) + ((loads - dem) * f32(c.max_load)) * f32(c.w)  # rule 6
    tc2 = D[:, 0, :] / f32(c.velocity)
    soc2 = _energy(c, mass2, G[:, 0, :], tc2)
    soc3 = soc0 + soc2
    soc3[:, :F + 1] = 0
    tc3 = (d0 + D[:, 0, :]) / f32(c.velocity)
    tc3[:, 1:F + 1] += f32(c.charging_time)
    tc3[:, F + 1:] += f32(c.serve_time)
    _, d3, s3, d4 = station_constants(c, D, G)                        # rule 7
    tc4 = d3 / f32(c.velocity)
    soc4 = _energy(c, mass2, s3, tc4)
    soc5 = soc0 + soc4
    tc5 = ((d0 + d3) + d4) / f32(c.velocity)
    tc5[:, F + 1:] += f32(c.charging_time + c.serve_time)
    m[(soc < soc0) | (tim < tc0)] = False
    m[((soc < soc3) | (tim < tc3)) & ((soc < soc5) | (tim < tc5))] = False
    all_masked = m[:, F + 1:].sum(1) <= 0                             # rule 8
    m[all_masked, 0] = True
    return m.astype(f32)


# ----------------------------------------------------------------------------------------------
# policy network
# ----------------------------------------------------------------------------------------------
def params_from_state_dict(sd):
    """state_dict (torch tensors or arrays) -> {name: fp32 ndarray}."""
    out = {}
    for k, v in sd.items():
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def _linear(x, W, b=None):
    y = x @ W.T.astype(f32)
    if b is not None:
        y = y + b
    return y.astype(f32)


def init_embed(P, static, dynamic, F):
    """nets/DRLModel.py:192-200 on information[:,:, [0,1,3]] = (x, y, demand)."""
    xy = (np.transpose(static, (0, 2, 1)) / f32(100)).astype(f32)          # [B,S,2]
    dem = dynamic[:, 1, :, None].astype(f32)
    e0 = _linear(xy[:, 0:1], P["init_embed_depot_and_station.weight"], P["init_embed_depot_and_station.bias"])
    e1 = _linear(xy[:, 1:F + 1], P["init_embed_station.weight"], P["init_embed_station.bias"])
    e2 = _linear(np.concatenate([xy[:, F + 1:], dem[:, F + 1:]], 2), P["init_embed.weight"], P["init_embed.bias"])
    return np.concatenate([e0, e1, e2], 1).astype(f32)


def _softmax(x, axis=-1):
    m = x.max(axis, keepdims=True)
    e = np.exp(x - m)
    return (e / e.sum(axis, keepdims=True)).astype(f32)


def _bn_eval(P, prefix, x, eps=1e-5):
    rm, rv = P[prefix + ".running_mean"], P[prefix + ".running_var"]
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    return ((x - rm) / np.sqrt(rv + f32(eps)) * w + b).astype(f32)


def _bn_train(P, prefix, x, eps=1e-5):
    flat = x.reshape(-1, x.shape[-1]).astype(np.float64)
    mu = flat.mean(0); var = flat.var(0)
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    return (((x - mu) / np.sqrt(var + eps)) * w + b).astype(f32)


def encoder(P, h, n_layers=3, n_heads=8, train_bn=False):
    """nets/GraphEncoder.py:202-214.  h [B,S,128] -> [B,S,128]."""
    B, S, d = h.shape
    bn = _bn_train if train_bn else _bn_eval
    for l in range(n_layers):
        pre = f"embedder.layers.{l}"
        Wq, Wk, Wv, Wo = (P[f"{pre}.0.module.{n}"] for n in ("W_query", "W_key", "W_val", "W_out"))
        flat = h.reshape(B * S, d)
        Q = np.matmul(flat[None], Wq).reshape(n_heads, B, S, -1)
        K = np.matmul(flat[None], Wk).reshape(n_heads, B, S, -1)
        V = np.matmul(flat[None], Wv).reshape(n_heads, B, S, -1)
        compat = f32(1.0 / math.sqrt(Wq.shape[-1])) * np.matmul(Q, np.swapaxes(K, 2, 3))
        attn = _softmax(compat)
        heads = np.matmul(attn, V)                                  # [H,B,S,dv]
        out = heads.transpose(1, 2, 0, 3).reshape(B * S, -1) @ Wo.reshape(-1, d)
        h = h + out.reshape(B, S, d)
        h = bn(P, f"{pre}.1.normalizer", h)
        ff = _linear(np.maximum(_linear(h, P[f"{pre}.2.module.0.weight"], P[f"{pre}.2.module.0.bias"]), 0),
                     P[f"{pre}.2.module.2.weight"], P[f"{pre}.2.module.2.bias"])
        h = bn(P, f"{pre}.3.normalizer", (h + ff).astype(f32))
    return h.astype(f32)


def precompute(P, emb, n_heads=8):
    """nets/DRLModel.py:345-362: fixed_ctx [B,128], gK,gV [B,H,S,16], lK [B,S,128]."""
    B, S, d = emb.shape
    fixed_ctx = _linear(emb.mean(1, dtype=f32), P["project_fixed_context.weight"])
    proj = _linear(emb, P["project_node_embeddings.weight"])
    gK, gV, lK = proj[..., :d], proj[..., d:2 * d], proj[..., 2 * d:]
    gK = gK.reshape(B, S, n_heads, -1).transpose(0, 2, 1, 3)
    gV = gV.reshape(B, S, n_heads, -1).transpose(0, 2, 1, 3)
    return fixed_ctx, np.ascontiguousarray(gK), np.ascontiguousarray(gV), np.ascontiguousarray(lK)


def route_context(P, c, dynamic, run_max):
    """nets/DRLModel.py:203-237.  run_max = elementwise max of the embeddings of every node picked
    so far (zeros before the first pick)."""
    veh = np.stack([dynamic[:, 0, 0], dynamic[:, 2, 0] / f32(c.Start_SOC), dynamic[:, 3, 0] / f32(c.t_limit)], 1)
    v1 = _linear(veh.astype(f32), P["tour.weight"])
    cat = np.concatenate([v1, run_max], 1)
    h1 = np.maximum(_linear(cat, P["FF_tour.0.weight"], P["FF_tour.0.bias"]), 0)
    return _linear(h1, P["FF_tour.2.weight"], P["FF_tour.2.bias"])


def step_context_am(P, emb, dynamic, now):
    """nets/AM.py:333-377: project_step_context([emb[now], load]) -- the AM variant's query term (no route context)."""
    B = emb.shape[0]
    cat = np.concatenate([emb[np.arange(B), now], dynamic[:, 0, 0:1]], 1).astype(f32)
    return _linear(cat, P["project_step_context.weight"])


def step_log_p(P, fixed, emb, now, mask, ctx, tanh_clipping=10., temp=1.0, n_heads=8, add_now=True):
    """nets/DRLModel.py:379-452 -> log_p [B,S].  add_now=False: nets/AM.py:336-337 (query = fixed context + ctx)."""
    fixed_ctx, gK, gV, lK = fixed
    B, S, d = emb.shape
    q = (fixed_ctx + ctx) + emb[np.arange(B), now] if add_now else fixed_ctx + ctx
    qh = q.reshape(B, n_heads, 1, -1)
    compat = np.matmul(qh, np.swapaxes(gK, 2, 3))[:, :, 0, :] / f32(math.sqrt(d // n_heads))   # [B,H,S]
    infeasible = (mask == 0)
    compat = np.where(infeasible[:, None, :], -np.inf, compat).astype(f32)
    attn = _softmax(compat)
    heads = np.matmul(attn[:, :, None, :], gV)[:, :, 0, :]          # [B,H,16]
    glimpse = _linear(heads.reshape(B, d), P["project_out.weight"])
    logits = np.matmul(lK, glimpse[:, :, None])[:, :, 0] / f32(math.sqrt(d))
    if tanh_clipping > 0:
        logits = np.tanh(logits) * f32(tanh_clipping)
    logits = np.where(infeasible, -np.inf, logits).astype(f32)
    z = logits / f32(temp)
    mx = z.max(1, keepdims=True)
    lse = np.log(np.exp(z - mx).sum(1, keepdims=True, dtype=f32)) + mx
    return (z - lse).astype(f32)


def rollout(P, static, dynamic, D, G, consts=None, decode_type="greedy", forced=None, rng=None,
            objective="EM", trace=False, train_bn=False, max_steps=1000, temp=1.0):
    """The whole forward of nets/DRLModel.py:109-140.

    Returns dict(pi [B,T] int64, ll [B,T], R [B,T], E [B,T] (energy), log_p [B,T,S] if trace, ...).
    ``forced`` [B,T] teacher-forces the actions (log_p is still the policy's)."""
    c = consts or EVRPConsts()
    static = static.astype(f32); dynamic = dynamic.astype(f32)
    B, _, S = static.shape
    emb = encoder(P, init_embed(P, static, dynamic, c.F), train_bn=train_bn)
    fixed = precompute(P, emb)
    mask = np.ones((B, S), f32)
    now = np.zeros(B, np.int64)
    run_max = np.zeros((B, emb.shape[2]), f32)
    am = "project_step_context.weight" in P
    pis, lls, Rs, Es, logps, masks, dyns = [], [], [], [], [], [], []
    ar = np.arange(B)
    for t in range(max_steps):
        if forced is not None:
            if t >= forced.shape[1]:
                break
        elif not mask.any():
            break
        if am:                                              # nets/AM.py variant
            log_p = step_log_p(P, fixed, emb, now, mask, step_context_am(P, emb, dynamic, now), temp=temp, add_now=False)
        else:
            ctx = route_context(P, c, dynamic, run_max)
            log_p = step_log_p(P, fixed, emb, now, mask, ctx, temp=temp)
        if forced is not None:
            sel = forced[:, t].astype(np.int64)
        elif decode_type == "greedy":
            sel = np.argmax(np.exp(log_p), 1)
        else:
            p = np.exp(log_p.astype(np.float64))
            cdf = np.cumsum(p, 1)
            u = rng.random_sample(B)[:, None] * cdf[:, -1:]
            sel = np.minimum((cdf < u).sum(1), S - 1)
            bad = mask[ar, sel] == 0
            if bad.any():                                   # numerical edge: fall back to the mode
                sel[bad] = np.argmax(log_p[bad], 1)
        if trace:
            masks.append(mask.copy()); logps.append(log_p)
        dynamic, e, dist = update_dynamic(c, dynamic, D, G, now, sel)
        mask = update_mask(c, dynamic, D, G, sel)
        if trace:
            dyns.append(dynamic.copy())
        run_max = emb[ar, sel] if t == 0 else np.maximum(run_max, emb[ar, sel])
        now = sel
        pis.append(sel); lls.append(log_p[ar, sel]); Es.append(e[:, 0])
        Rs.append(e[:, 0] if objective == "EM" else dist[:, 0])
    out = dict(pi=np.stack(pis, 1).astype(np.int64), ll=np.stack(lls, 1).astype(f32),
               R=np.stack(Rs, 1).astype(f32), E=np.stack(Es, 1).astype(f32), emb=emb, fixed=fixed)
    if trace:
        out.update(log_p=np.stack(logps, 1), mask=np.stack(masks, 1), dynamic=np.stack(dyns, 1))
    return out


def sample_many(P, static, dynamic, D, G, batch_rep, iter_rep=1, consts=None, rng=None):
    """nets/DRLModel.py:283-313 -> (mincosts [B], minpis [B,Tmax])."""
    rng = rng or np.random.RandomState(0)
    B = static.shape[0]
    rep = lambda a: np.broadcast_to(a[None], (batch_rep,) + a.shape).reshape((-1,) + a.shape[1:])
    costs, pis = [], []
    for _ in range(iter_rep):
        o = rollout(P, rep(static), rep(dynamic), rep(D), rep(G), consts, "sample", rng=rng)
        costs.append(o["R"].sum(1).reshape(batch_rep, B).T)
        pis.append(o["pi"].reshape(batch_rep, B, -1).transpose(1, 0, 2))
    costs = np.concatenate(costs, 1)
    T = max(p.shape[-1] for p in pis)
    pis = np.concatenate([np.pad(p, ((0, 0), (0, 0), (0, T - p.shape[-1]))) for p in pis], 1)
    am = costs.argmin(1)
    return costs[np.arange(B), am], pis[np.arange(B), am]


def reinforce_loss(ll, R, baseline):
    """trainer.py:116-121."""
    reward = R.sum(1)
    adv = reward - baseline
    return float((adv * ll.sum(1)).mean()), adv
