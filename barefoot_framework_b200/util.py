"""Host-side mirror of the reference's util.py for the hot path: same function names, argument
meaning and return shapes, with the numerics routed to the sm_100a kernels (engine.py).

Where the reference maps ONE work item per (model jj, fantasy point kk, hyper-parameter set mm)
through a process pool (barefoot.py:1070-1073, 1271-1283, 2001-2013), the ``*_batch`` functions
here take the whole grid in one call: all sets are factorised together and swept by one
posterior + acquisition launch.  The per-item functions (``calculate_KG`` ... ``fused_calculate``,
``batchAcquisitionFunc``, ``evaluateFusedModel``) keep the reference's pickle-file signature for
drop-in use and call the batch form with a grid of one."""
from pickle import dump, load

import numpy as np
import torch

from . import dist as bdist
from . import engine
from .acquisitionFunc import thompson_sampling
from .engine import ACQ_BITS, ACQ_EI, ACQ_GREEDY, ACQ_PI, ACQ_UCB, SLOT_EI, SLOT_GREEDY, SLOT_PI, SLOT_UCB

XI_ROM = 0.01            # util.py:323-326 (calculate_EI), :449, :919
SN_KG = 0.1              # util.py:260,619
_SLOTS = {"EI": SLOT_EI, "PI": SLOT_PI, "UCB": SLOT_UCB, "Greedy": SLOT_GREEDY}


# ---------------------------------------------------------------------------------------------
# batch selection
# ---------------------------------------------------------------------------------------------
def kmedoids_max(input_arr, n_clust):
    """util.py:38-49: cluster with sklearn_extra KMedoids(n_clust, random_state=0) defaults and
    return, per cluster, the index of the first row of the WHOLE array whose column 0 equals the
    cluster's maximum column-0 value (quirk Q14).  Distances, labels and medoid updates run on
    the device (bf_kmedoids_*); the per-cluster maximum is a host pass over N labels."""
    input_arr = np.ascontiguousarray(np.asarray(input_arr, dtype=np.float64))
    _, labels, _ = engine.kmedoids_fit(input_arr, int(n_clust))
    new_medoids = []
    col0 = input_arr[:, 0]
    for ii in range(n_clust):
        members = col0[labels == ii]
        if members.shape[0] == 0:          # reference: np.max of an empty array -> ValueError -> pass
            continue
        new_medoids.append(np.where(col0 == np.max(members))[0][0])
    return np.array(new_medoids)


def call_model(param):
    """util.py:51-55."""
    return param["Model"](param["Input Values"])


def storeObject(obj, filename):
    with open(filename, 'wb') as f:
        dump(obj, f)


def checkInitInputs(*args, **kwargs):
    pass


def checkParameterInputs(*args, **kwargs):
    pass


# ---------------------------------------------------------------------------------------------
# one sweep = posterior + acquisition(s) for every set
# ---------------------------------------------------------------------------------------------
def _ts_discard(n_sets, n_candidates):
    """Advance the global NumPy RNG exactly as thompson_sampling would for n_sets sets it does not evaluate here
    (one normal per candidate and set, acquisitionFunc.py:243): ranks that shard the sets keep ONE random stream -
    every rank ends in the state a single process would have, and each set sees the draws it would have seen."""
    for _ in range(int(n_sets)):
        np.random.normal(size=int(n_candidates))


def _sweep(f, scale, shift, x_test, names, stage, curr_max, iteration, xi, M=None):
    """Run the fused posterior for every set of `f` and return {name: (values [S], indices [S])}
    as host arrays, plus "max_mean" [S].  PI / UCB spread: sqrt(var) in the ROM stage
    (util.py:451,519), var in the TM / batch / Hedge paths (util.py:598-612,1058-1072,764-789)."""
    names = list(names)
    need_kg = "KG" in names
    need_ts = "TS" in names
    mask = ACQ_GREEDY
    for nm in names:
        mask |= ACQ_BITS.get(nm, 0)
    d = f.d
    kt = engine.ucb_kt(d, iteration) if "UCB" in names else 0.0
    res = engine.posterior_sweep(f, x_test, scale, shift, want_mu=need_kg or need_ts,
                                 want_var=need_kg or need_ts, acq_mask=mask,
                                 spread_is_sqrt=(stage == "ROM"), curr_max=curr_max, xi=xi, kt=kt)
    out = {}
    if need_kg:
        v, i, _ = engine.kg_from_sweep(res, M=M, sn=SN_KG)
        out["KG"] = (v.cpu().numpy(), i.cpu().numpy())
    bv, bi = res.best_val.cpu().numpy(), res.best_idx.cpu().numpy()
    for nm in names:
        if nm in _SLOTS:
            out[nm] = (bv[:, _SLOTS[nm]].copy(), bi[:, _SLOTS[nm]].copy())
    out["max_mean"] = bv[:, SLOT_GREEDY].copy()
    if need_ts:
        # Thompson sampling draws from the host RNG per set, in set order, like the reference's
        # work items would in a serial pool (acquisitionFunc.py:216-248)
        mu, var = res.mu.cpu().numpy(), res.var.cpu().numpy()
        tv, ti = np.empty(f.S), np.empty(f.S, dtype=np.int64)
        for s in range(f.S):
            tv[s], ti[s], _ = thompson_sampling(mu[s], np.sqrt(var[s]))
        out["TS"] = (tv, ti)
    return out


def _sweep_candidate_sharded(f, scale, shift, x_test, name, stage, curr_max, iteration, xi, M=None):
    """_sweep for ONE deterministic acquisition when there are fewer sets than ranks (reification-only runs have
    H = 1): the CANDIDATES are split over the ranks instead (SURVEY.md 8e "secondary: candidate tiles"), every rank
    sweeps its tile for all sets, and the per-rank records are combined with numpy's first-index rule
    (dist.reduce_candidate_shards).  KG needs the top two fused means over ALL candidates first
    (dist.reduce_top2), then the closed form runs locally.  Returns (values [S], global indices [S]) on the host."""
    N = x_test.shape[0]
    lo, hi = bdist.shard_range(N)
    kg = name == "KG"
    mask = ACQ_GREEDY | ACQ_BITS.get(name, 0)
    kt = engine.ucb_kt(f.d, iteration) if name == "UCB" else 0.0
    res = engine.posterior_sweep(f, x_test[lo:hi], scale, shift, want_mu=kg, want_var=kg, acq_mask=mask,
                                 spread_is_sqrt=(stage == "ROM"), curr_max=curr_max, xi=xi, kt=kt)
    if not kg:
        slot = _SLOTS[name]
        v, i = bdist.reduce_candidate_shards(res.best_val[:, slot].contiguous(), res.best_idx[:, slot].contiguous(), lo)
        return v.cpu().numpy(), i.cpu().numpy()
    t1, gi1, t2 = bdist.reduce_top2(res.best_val[:, SLOT_GREEDY].contiguous(), res.best_idx[:, SLOT_GREEDY].contiguous(),
                                    res.best_val[:, engine.SLOT_TOP2].contiguous(), lo)
    res.best_val[:, SLOT_GREEDY], res.best_val[:, engine.SLOT_TOP2] = t1, t2
    res.best_idx[:, SLOT_GREEDY] = gi1 - lo                     # local position of the global arg-max (or outside the tile)
    m_total = N if M is None else int(M)
    v, i, _ = engine.kg_from_sweep(res, M=max(0, min(hi, m_total) - lo), sn=SN_KG)
    v, i = bdist.reduce_candidate_shards(v, i, lo)
    i = torch.where(v > 0, i, torch.zeros_like(i))              # nothing above the reference's floor: (0, index 0)
    return v.cpu().numpy(), i.cpu().numpy()


_HEDGE_ORDER = ("KG", "TS", "EI", "Greedy", "PI", "UCB")      # util.py:626-654, 733-789


# ---------------------------------------------------------------------------------------------
# TM stage: fused_calculate for all hyper-parameter sets (util.py:543-669)
# ---------------------------------------------------------------------------------------------
def fused_calculate_batch(model, x_fused, hp_sets, kernel, x_test, curr_max, sampleOpt,
                          iteration=1, xi=0.01, targets=None):
    """Every work item of barefoot.py:1222-1283 in one device pass.  Returns (values [H],
    indices [H]); for sampleOpt "Hedge" a list of H lists of six [value, index] pairs."""
    x_test = np.asarray(x_test, dtype=np.float64)
    hp = np.atleast_2d(np.asarray(hp_sets, dtype=np.float64))
    deterministic = sampleOpt not in ("Hedge", "TS")          # host RNG draws must stay in one stream
    name = sampleOpt if sampleOpt in ("KG", "TS", "EI", "PI", "UCB") else "Greedy"
    world = bdist.world()[1]
    if deterministic and world > 1 and hp.shape[0] < world <= x_test.shape[0]:
        # fewer sets than ranks (H = 1 in reification-only runs): split the candidates instead
        err = out = None
        try:
            f, scale, shift = model.create_fused_GP_batch(x_fused, hp, kernel, targets=targets)
        except Exception as e:                                 # noqa: BLE001 - identical on every rank (same sets)
            err = e
        bdist.raise_together(err, "fused_calculate_batch")
        return _sweep_candidate_sharded(f, scale, shift, x_test, name, "TM", curr_max, iteration, xi)
    # hyper-parameter sets are sharded over the ranks of a torchrun job; only (value, index) records travel.  The
    # host-RNG members (TS, and TS inside Hedge) are sharded too: a rank draws for its own sets and discards the
    # draws of the others in set order (_ts_discard), so the stream is the single-process one on every rank.
    H_all, N_all = hp.shape[0], x_test.shape[0]
    lo, hi = bdist.shard_range(H_all)
    names = _HEDGE_ORDER if sampleOpt == "Hedge" else [name]
    err = r = f = None
    try:
        f, scale, shift = model.create_fused_GP_batch(x_fused, hp[lo:hi], kernel, targets=targets)
        if not deterministic:
            _ts_discard(lo, N_all)
        r = _sweep(f, scale, shift, x_test, names, "TM", curr_max, iteration, xi)
        if not deterministic:
            _ts_discard(H_all - hi, N_all)
    except Exception as e:                                     # noqa: BLE001 - re-raised on every rank below
        err = e
    # a set that is not positive definite (or any device error) on ONE rank must fail the stage on all
    bdist.raise_together(err, "fused_calculate_batch")
    g = {nm: bdist.gather_host_sets(*r[nm], n_total=H_all) for nm in names}
    if sampleOpt == "Hedge":
        return [[[g[nm][0][h], int(g[nm][1][h])] for nm in _HEDGE_ORDER] for h in range(H_all)]
    return g[name]


def fused_calculate(param):
    """util.py:543-669 with the reference's pickle-file work-item signature."""
    with open("data/parameterSets/parameterSet{}".format(param[1]), 'rb') as fh:
        data = load(fh)
    with open("data/reificationObj", 'rb') as fh:
        model_temp = load(fh)
    (iteration, model_data, x_fused, fused_model_HP, kernel, x_test, curr_max, xi, sampleOpt) = data[param[0]]
    out = fused_calculate_batch(model_temp, x_fused, np.asarray(fused_model_HP)[None], kernel, x_test,
                                curr_max, sampleOpt, iteration, xi)
    if sampleOpt == "Hedge":
        return out[0], x_test.shape[0]
    return [out[0][0], int(out[1][0])], x_test.shape[0]


# ---------------------------------------------------------------------------------------------
# ROM stage: calculate_* for the (jj, kk, mm) grid (util.py:216-886)
# ---------------------------------------------------------------------------------------------
def rom_stage_batch(model, x_fused, hp_sets, kernel, x_test, new_mean, acq, cost, curr_max,
                    iteration=1, true_sample_count=None, jj_list=None, kk_list=None, as_arrays=False,
                    max_bytes_in_flight=16 << 30, solve_path=None, fantasy="append"):
    """The ROM-stage grid of barefoot.py:1001-1073: for every model jj, fantasy point kk and
    hyper-parameter set mm, update ROM jj with (x_test[kk], new_mean[jj][kk]), rebuild the fused
    GP and evaluate `acq` over x_test.  Returns the rows [max fused mean, value / cost[jj], x*,
    jj, kk, mm] in the reference's loop order (Greedy is not divided by the cost, util.py:876).
    For "Hedge" returns six such row lists in the portfolio order [KG, TS, EI, Greedy, PI, UCB].
    as_arrays=True returns float64 arrays [G * H, 6 (+ 2 d for Greedy)] instead of lists of rows.
    solve_path: None = automatic (few candidates per fused GP -> engine.SolveSpec, else explicit factors).
    fantasy: "append" (default) = all fantasy updates of a model at once through the bordered form of the
    appended Cholesky factor (model_reification.fantasy_targets_batched); "refactor" = the reference's literal
    per-group update_GP (append + full refactor) in a host loop - the same numbers to rounding.

    The fused targets of one (jj, kk) do not depend on mm; the H fused GPs of every (jj, kk)
    are factorised and swept together (one set per (jj, kk, mm))."""
    import copy
    x_test = np.asarray(x_test, dtype=np.float64)
    hp = np.atleast_2d(np.asarray(hp_sets, dtype=np.float64))
    H = hp.shape[0]
    N = x_test.shape[0]
    M = N if true_sample_count is None else int(true_sample_count)
    jj_list = list(range(model.num_models)) if jj_list is None else list(jj_list)
    kk_list = list(range(N)) if kk_list is None else list(kk_list)
    hedge = acq == "Hedge"
    names = _HEDGE_ORDER if hedge else [acq if acq in ("KG", "TS", "EI", "PI", "UCB") else "Greedy"]
    all_keys = [(jj, kk) for jj in jj_list for kk in kk_list]
    # under torchrun the (model, fantasy point) groups are split over the ranks (all H sets of a group stay
    # local) and the finished rows are gathered in group order; for the host-RNG members (TS, Hedge) a rank
    # discards the draws of the groups it does not hold, in order, so that the stream stays the single-process one
    sharded = bdist.world()[1] > 1
    uses_rng = acq in ("Hedge", "TS")
    g_lo, g_hi = bdist.shard_range(len(all_keys)) if sharded else (0, len(all_keys))
    keys = all_keys[g_lo:g_hi]
    G = len(keys)
    # The factors of one set take ~8 F^2 / 2 bytes (1 MB at F = 500): the (jj, kk) groups are processed in
    # chunks of at most `max_sets_in_flight` sets so the literal ROM stage of the reference (3 models x 50
    # candidates x 1000 hyper-parameter sets = 150000 fused GPs, 159 GB of factors) fits any budget.
    F = np.asarray(x_fused).shape[0]
    use_solve = solve_path if solve_path is not None else engine.SolveSpec.supported(F, N)
    per_set = 8 * (3 * F + N) if use_solve else 8 * (engine.linv_packed_doubles(F) + 3 * F)
    groups_per_chunk = max(1, int(max_bytes_in_flight // (per_set * H)))
    parts = {nm: ([], []) for nm in names}
    parts["max_mean"] = []
    err = None
    try:
        base_means, base_vars = model.fused_rows_device(x_fused)
        if uses_rng:
            _ts_discard(g_lo * H, N)
        for g0 in range(0, G, groups_per_chunk):
            chunk = keys[g0:g0 + groups_per_chunk]
            if fantasy == "refactor":
                zs, zerrs, stats = [], [], []
                for (jj, kk) in chunk:
                    # the reference's literal step: a private copy of the model, update_GP (append + full refactor)
                    m2 = _fantasy_copy(model, copy)
                    m2.update_GP(np.expand_dims(x_test[kk], axis=0), np.expand_dims(np.array([new_mean[jj][kk]]), axis=0), jj)
                    # only ROM jj changed: its row is recomputed, the other rows (and every discrepancy GP) are the
                    # unmodified model's (util.py:243-246 recomputes all 2m posteriors per work item)
                    mj, vj = m2.fused_rows_device(x_fused, models=[jj])
                    means, variances = base_means.clone(), base_vars.clone()
                    means[jj], variances[jj] = mj[0], vj[0]
                    z, zerr, st = m2.fused_targets_from_rows(means, variances)
                    zs.append(z)
                    zerrs.append(zerr)
                    stats.append(st)
                Zg, ZEg, ST = torch.stack(zs), torch.stack(zerrs), torch.stack(stats)
            else:
                # all fantasy points of a model in one launch sequence (bordered append, no per-group host work)
                zs, zerrs, stats = [], [], []
                a = 0
                while a < len(chunk):
                    jj = chunk[a][0]
                    b = a
                    while b < len(chunk) and chunk[b][0] == jj:
                        b += 1
                    kks = [kk for _, kk in chunk[a:b]]
                    z, zerr, st = model.fantasy_targets_batched(x_fused, x_test[kks], [new_mean[jj][kk] for kk in kks],
                                                                jj, base_means, base_vars)
                    zs.append(z)
                    zerrs.append(zerr)
                    stats.append(st)
                    a = b
                Zg, ZEg, ST = torch.cat(zs), torch.cat(zerrs), torch.cat(stats)
            Gc = len(chunk)
            Z = Zg.repeat_interleave(H, dim=0).contiguous()                       # [Gc*H, F]
            ZE = ZEg.repeat_interleave(H, dim=0).contiguous()                     # ST: [Gc, 2] (mean, std)
            scale = ST[:, 1].repeat_interleave(H).contiguous()
            shift = ST[:, 0].repeat_interleave(H).contiguous()
            set_hp = np.tile(np.arange(H, dtype=np.int32), Gc)
            if use_solve:
                # few candidates per fused GP: factor + forward substitution, no explicit inverse (engine.SolveSpec)
                f = engine.SolveSpec(kernel, x_fused, Z, hp[:, 1:] ** 2, hp[:, 0], ZE, ndim=model.num_dim, set_hp=set_hp)
            else:
                f = engine.build_factors(kernel, x_fused, Z, hp[:, 1:] ** 2, hp[:, 0], ZE, ndim=model.num_dim,
                                         set_hp=set_hp)
            # calculate_GPHedge hands the variance to PI / UCB (util.py:764-789) and uses x_test.shape[0] for KG
            rc = _sweep(f, scale, shift, x_test, names, "TM" if hedge else "ROM", curr_max, iteration, XI_ROM,
                        M=(N if hedge else M))
            for nm in names:
                parts[nm][0].append(np.asarray(rc[nm][0], dtype=np.float64))
                parts[nm][1].append(np.asarray(rc[nm][1], dtype=np.int64))
            parts["max_mean"].append(np.asarray(rc["max_mean"], dtype=np.float64))
            del f, Z, ZE
        if uses_rng:
            _ts_discard((len(all_keys) - g_hi) * H, N)
    except Exception as e:                                     # noqa: BLE001 - re-raised below on every rank
        err = e
    if sharded:
        bdist.raise_together(err, "rom_stage_batch")      # before the row gather: all ranks raise or none
    elif err is not None:
        raise err
    r = {nm: (np.concatenate(parts[nm][0]) if G else np.zeros(0), np.concatenate(parts[nm][1]) if G else np.zeros(0, dtype=np.int64))
         for nm in names}
    r["max_mean"] = np.concatenate(parts["max_mean"]) if G else np.zeros(0)
    # rows [max fused mean, value / cost[jj], x*, jj, kk, mm] for every set, built as arrays (no Python
    # loop over the G * H work items); set s = g * H + mm
    jj_s = np.repeat(np.array([k_[0] for k_ in keys], dtype=np.int64), H)
    kk_s = np.repeat(np.array([k_[1] for k_ in keys], dtype=np.int64), H)
    mm_s = np.tile(np.arange(H, dtype=np.int64), G)
    cost_s = np.asarray(cost, dtype=np.float64)[jj_s]
    rows = {}
    for nm in names:
        val = np.asarray(r[nm][0], dtype=np.float64)
        idx = np.asarray(r[nm][1], dtype=np.int64)
        if not (nm == "Greedy" and not hedge):
            val = val / cost_s
        cols = [np.asarray(r["max_mean"], dtype=np.float64), val, idx.astype(np.float64), jj_s.astype(np.float64),
                kk_s.astype(np.float64), mm_s.astype(np.float64)]
        if nm == "Greedy" and not hedge:
            # calculate_Greedy appends the coordinates of the maximum twice (util.py:878-884)
            cols += [x_test[idx], x_test[idx]]
        rows[nm] = np.column_stack(cols)
        if sharded:
            w_ = bdist.world()[1]
            per_rank = [(b_ - a_) * H for a_, b_ in (bdist.shard_range(len(all_keys), q, w_) for q in range(w_))]
            rows[nm] = bdist.gather_row_blocks(engine.to_dev(np.ascontiguousarray(rows[nm])), per_rank).cpu().numpy()
    if as_arrays:
        return [rows[nm] for nm in names] if hedge else rows[names[0]]

    def as_lists(a):
        # the reference's work items return Python lists with int indices (util.py:241,265-266)
        return [[row[0], row[1], int(row[2]), int(row[3]), int(row[4]), int(row[5])] + list(row[6:]) for row in a.tolist()]
    if hedge:
        return [as_lists(rows[nm]) for nm in names]
    return as_lists(rows[names[0]])


def _fantasy_copy(model, copy):
    """A model_reification whose ROM GPs can be updated without touching `model` (the reference
    gets this from unpickling data/reificationObj per work item, util.py:236-237)."""
    m2 = copy.copy(model)
    m2.x_train = [np.array(x) for x in model.x_train]
    m2.y_train = [np.array(y) for y in model.y_train]
    m2.gp_models = [copy.copy(g) for g in model.gp_models]
    return m2


def _rom_item(param, acq):
    with open("data/parameterSets/parameterSet{}".format(param[1]), 'rb') as fh:
        data = load(fh)
    with open("data/reificationObj", 'rb') as fh:
        model_temp = load(fh)
    (iteration, model_data, x_fused, fused_model_HP, kernel, x_test, jj, kk, mm, true_sample_count,
     cost, curr_max) = data[param[0]]
    new_mean = {jj: {kk: float(np.asarray(model_data[1]).reshape(-1)[0])}}
    x_test = np.asarray(x_test, dtype=np.float64)
    # the fantasy point travels in model_data; it is x_test[kk] in every reference call site
    xt = x_test.copy()
    xt_kk = np.asarray(model_data[0], dtype=np.float64).reshape(-1)
    if not np.array_equal(xt[kk], xt_kk):
        raise ValueError("model_data[0] must be x_test[kk] (barefoot.py:1005-1007)")
    out = rom_stage_batch(model_temp, x_fused, np.asarray(fused_model_HP)[None], kernel, xt, new_mean, acq,
                          cost, curr_max, iteration, true_sample_count, jj_list=[jj], kk_list=[kk])
    if acq == "Hedge":
        return [[o[0][0], o[0][1], o[0][2], jj, kk, mm] for o in out]
    row = out[0]
    return [row[0], row[1], row[2], jj, kk, mm] + list(row[6:])


def calculate_KG(param):
    """util.py:216-277."""
    return _rom_item(param, "KG")


def calculate_EI(param):
    """util.py:279-341."""
    return _rom_item(param, "EI")


def calculate_TS(param):
    """util.py:345-405."""
    return _rom_item(param, "TS")


def calculate_PI(param):
    """util.py:407-469."""
    return _rom_item(param, "PI")


def calculate_UCB(param):
    """util.py:471-541."""
    return _rom_item(param, "UCB")


def calculate_GPHedge(param):
    """util.py:671-823."""
    return _rom_item(param, "Hedge")


def calculate_Greedy(param):
    """util.py:825-886."""
    return _rom_item(param, "Greedy")


# ---------------------------------------------------------------------------------------------
# batch-only path: batchAcquisitionFunc for all hyper-parameter sets (util.py:913-1129)
# ---------------------------------------------------------------------------------------------
def batch_acquisition_batch(modelGP, hp_sets, x_test, curr_max, acqFunc, iteration=1, xi=0.01):
    """Every work item of barefoot.py:1966-2013 (reification=False) in one device pass.  The
    hyper-parameters overwrite l_param WITHOUT squaring (util.py:948, quirk Q1).  Returns
    (values [H], indices [H]), or the Hedge list of lists.  KG on this path uses the FULL
    posterior covariance (util.py:1078-1081) and runs per set through bf_kg_full."""
    x_test = np.asarray(x_test, dtype=np.float64)
    hp = np.atleast_2d(np.asarray(hp_sets, dtype=np.float64))
    H_total = hp.shape[0]
    N_all = x_test.shape[0]
    world = bdist.world()[1]
    by_candidates = acqFunc not in ("KG", "Hedge", "TS") and world > 1 and H_total < world <= N_all
    x = np.asarray(modelGP.x_train, dtype=np.float64)
    x = x.reshape(x.shape[0], -1)
    y = np.asarray(modelGP.y_train, dtype=np.float64) - float(modelGP.mean)
    if by_candidates:
        # fewer sets than ranks: every rank factorises all sets and sweeps ITS candidate tile
        err = None
        try:
            f = engine.build_factors(modelGP.kern, x, y, hp[:, 1:], hp[:, 0], modelGP.sigma_n, ndim=modelGP.n_dim)
        except Exception as e:                                 # noqa: BLE001 - identical on every rank
            err = e
        bdist.raise_together(err, "batch_acquisition_batch")
        shift = torch.full((f.S,), float(modelGP.mean), dtype=engine.F64, device=f.X.device) if modelGP.mean else None
        nm = acqFunc if acqFunc in ("EI", "PI", "UCB") else "Greedy"
        return _sweep_candidate_sharded(f, None if shift is None else torch.ones_like(shift), shift, x_test, nm, "BATCH",
                                        curr_max, iteration, xi)
    # sets sharded over the ranks and gathered below; the host-RNG members draw for their own sets and discard the
    # other sets' draws in set order (_ts_discard)
    lo, hi = bdist.shard_range(H_total)
    hp = hp[lo:hi]
    uses_rng = acqFunc in ("Hedge", "TS")
    name = acqFunc if acqFunc in ("TS", "EI", "PI", "UCB") else "Greedy"
    err = r = f = None
    try:
        f = engine.build_factors(modelGP.kern, x, y, hp[:, 1:], hp[:, 0], modelGP.sigma_n, ndim=modelGP.n_dim,
                                 keep_dense=acqFunc in ("KG", "Hedge"))
        shift = None
        if modelGP.mean:
            shift = torch.full((f.S,), float(modelGP.mean), dtype=engine.F64, device=f.X.device)
        scale = None if shift is None else torch.ones_like(shift)
        if uses_rng:
            _ts_discard(lo, N_all)
        if acqFunc in ("KG", "Hedge"):
            names = [n for n in (_HEDGE_ORDER if acqFunc == "Hedge" else ()) if n != "KG"]
            r = _sweep(f, scale, shift, x_test, names, "BATCH", curr_max, iteration, xi) if names else {}
            r["KG"] = engine.kg_full_from_factors(f, x_test, shift=float(modelGP.mean), sn=SN_KG) if f.S else \
                (np.zeros(0), np.zeros(0, dtype=np.int64))
        else:
            r = _sweep(f, scale, shift, x_test, [name], "BATCH", curr_max, iteration, xi)
        if uses_rng:
            _ts_discard(H_total - hi, N_all)
    except Exception as e:                                     # noqa: BLE001 - re-raised on every rank below
        err = e
    bdist.raise_together(err, "batch_acquisition_batch")       # before the gather: all ranks raise or none
    members = _HEDGE_ORDER if acqFunc == "Hedge" else (["KG"] if acqFunc == "KG" else [name])
    g = {nm: bdist.gather_host_sets(*r[nm], n_total=H_total) for nm in members}
    if acqFunc == "Hedge":
        return [[[g[nm][0][h], int(g[nm][1][h])] for nm in _HEDGE_ORDER] for h in range(H_total)]
    return g[members[0]]


def batch_acquisition_multi(modelGP, hp_sets, x_test, curr_max, names=("EI", "PI", "UCB"), iteration=1, xi=0.01,
                            return_device=False):
    """Several elementwise acquisitions of the batch-only work item (util.py:913-1129) from ONE posterior
    sweep: GP set-up for every hyper-parameter set, posterior over x_test (host array, pinned host tensor or
    device tensor), then every requested acquisition fused into the epilogue.  Returns
    {name: (values [H], indices [H])} as host arrays, or device tensors with return_device (plus "status" for
    check_status, see below)."""
    hp = np.atleast_2d(np.asarray(hp_sets, dtype=np.float64))
    # the candidates travel host -> device on a side stream while the GPs are factorised
    x_test, copied = engine.to_dev_async(x_test)
    x = np.asarray(modelGP.x_train, dtype=np.float64)
    x = x.reshape(x.shape[0], -1)
    y = np.asarray(modelGP.y_train, dtype=np.float64) - float(modelGP.mean)
    # the factorisation status is read after the sweep (where the results are read anyway), so the host
    # does not stall between the set-up and the sweep
    # ... on a side stream, so that a caller who pipelines work items gets this item's set-up under the previous
    # item's sweep (engine.build_factors_overlapped)
    f = engine.build_factors_overlapped(modelGP.kern, x, y, hp[:, 1:], hp[:, 0], modelGP.sigma_n, ndim=modelGP.n_dim,
                                        check=False)
    if copied is not None:
        torch.cuda.current_stream().wait_event(copied)
        x_test.record_stream(torch.cuda.current_stream())
    mask = 0
    for nm in names:
        mask |= ACQ_BITS[nm]
    shift = None
    if modelGP.mean:
        shift = torch.full((f.S,), float(modelGP.mean), dtype=engine.F64, device=f.X.device)
    res = engine.posterior_sweep(f, x_test, None if shift is None else torch.ones_like(shift), shift, acq_mask=mask,
                                 spread_is_sqrt=False, curr_max=curr_max, xi=xi,
                                 kt=engine.ucb_kt(f.d, iteration) if "UCB" in names else 0.0)
    if return_device:
        # nothing is read back here: the caller reads the records when it needs them and hands "status" (the
        # per-set factorisation status, 0 = ok) to check_status() at that point, so consecutive work items
        # pipeline (the next item's copy and set-up are queued while this sweep runs)
        out = {nm: (res.best_val[:, _SLOTS[nm]], res.best_idx[:, _SLOTS[nm]]) for nm in names}
        out["status"] = f.info
        return out
    f.raise_if_not_positive_definite()
    bv, bi = res.best_val.cpu().numpy(), res.best_idx.cpu().numpy()
    return {nm: (bv[:, _SLOTS[nm]].copy(), bi[:, _SLOTS[nm]].copy()) for nm in names}


def check_status(status):
    """Raise what george / SciPy raise from the Cholesky factorisation (gpModel.py:41) if any set of a
    return_device call was not positive definite; `status` is that call's "status" entry (device or host)."""
    bad = int(torch.as_tensor(status).max().item()) if len(status) else 0
    if bad:
        raise np.linalg.LinAlgError("%d-th leading minor of a GP kernel matrix is not positive definite" % bad)


def batchAcquisitionFunc(param):
    """util.py:913-1129 (single-objective branch) with the reference's work-item signature."""
    with open("data/parameterSets/parameterSet{}".format(param[1]), 'rb') as fh:
        data = load(fh)
    with open("data/reificationObj", 'rb') as fh:
        modelGP = load(fh)
    iteration, x_test, fusedModelHP, curr_max, acqFunc, extra = data[param[0]]
    if isinstance(modelGP, list):
        raise NotImplementedError("multi-objective EHVI is outside the B200 hot path (DESIGN.md section 7)")
    out = batch_acquisition_batch(modelGP, np.asarray(fusedModelHP)[None], x_test, curr_max, acqFunc, iteration)
    if acqFunc == "Hedge":
        return out[0]
    return [out[0][0], int(out[1][0])]


def evaluate_fused_model_batch(model, reification, x_fused, hp_sets, kernel, x_test):
    """evaluateFusedModel (util.py:889-911) for all sets: the fused (or single-GP) posterior
    mean at x_test, [H, N] on the host."""
    hp = np.atleast_2d(np.asarray(hp_sets, dtype=np.float64))
    if reification:
        f, scale, shift = model.create_fused_GP_batch(x_fused, hp, kernel)
    else:
        x = np.asarray(model.x_train, dtype=np.float64)
        x = x.reshape(x.shape[0], -1)
        f = engine.build_factors(model.kern, x, np.asarray(model.y_train, dtype=np.float64) - float(model.mean),
                                 hp[:, 1:], hp[:, 0], model.sigma_n, ndim=model.n_dim)
        scale = shift = None
    res = engine.posterior_sweep(f, np.asarray(x_test, dtype=np.float64), scale, shift, want_mu=True)
    mu = res.mu.cpu().numpy()
    if not reification and model.mean:
        mu = mu + float(model.mean)
    return mu


def evaluateFusedModel(param):
    """util.py:889-911."""
    with open("data/parameterSets/parameterSet{}".format(param[1]), 'rb') as fh:
        data = load(fh)
    with open("data/reificationObj", 'rb') as fh:
        model_temp = load(fh)
    (finish, reification, x_fused, fused_model_HP, kernel, x_test, curr_max, xi, acqIndex) = data[param[0]]
    mu = evaluate_fused_model_batch(model_temp, reification, x_fused, np.asarray(fused_model_HP)[None], kernel, x_test)
    return [acqIndex, mu[0]]


# ---------------------------------------------------------------------------------------------
# design-space sampling and constraints (host side, O(N d); util.py:57-214 of the reference)
# ---------------------------------------------------------------------------------------------
def lhs(n, samples=None):
    """pyDOE.lhs(n, samples) with criterion None ("classic"), on the GLOBAL legacy numpy RNG like
    pyDOE: one uniform draw per stratum and column (a single rand(samples, n) call), then an
    independent np.random.permutation per column.  pyDOE is an unpinned third-party dependency of
    the reference (util.py:9); this follows its published algorithm so seeded runs reproduce."""
    samples = n if samples is None else samples
    edges = np.linspace(0, 1, samples + 1)
    u = np.random.rand(samples, n)
    lo, hi = edges[:samples], edges[1:samples + 1]
    pts = u * (hi - lo)[:, None] + lo[:, None]
    out = np.empty_like(pts)
    for j in range(n):
        out[:, j] = pts[np.random.permutation(range(samples)), j]
    return out


def cartesian(*arrays):
    """util.py:57-65: all combinations of the per-dimension arrays, one column per dimension."""
    mesh = np.meshgrid(*arrays)
    return np.reshape(np.concatenate(mesh).ravel(), (len(mesh), mesh[0].size)).T


def composition_sampler(ndim, nsamples):
    """util.py:67-70."""
    X = -np.log(lhs(ndim + 1, nsamples))
    X = X / np.sum(X, axis=1)[:, None]
    return X[:, 0:ndim]


def sampleDesignSpace(ndim, nsamples, sampleScheme):
    """util.py:72-102: "LHS", "Grid" (smallest full factorial grid with at least nsamples points),
    "Custom" (rows of data/possibleInputs.csv topped up with LHS) or "CompFunc"."""
    import pandas as pd
    if sampleScheme == "LHS":
        x = lhs(ndim, nsamples)
    if sampleScheme == "Grid":
        for per_axis in range(1, nsamples):
            x = cartesian(*([np.linspace(0, 1, per_axis)] * ndim))
            if x.shape[0] >= nsamples:
                return x
    if sampleScheme == "Custom":
        dfInputs = pd.read_csv("data/possibleInputs.csv", index_col=0)
        if dfInputs.shape[0] > nsamples:
            x = dfInputs.sample(n=nsamples)
        else:
            extra = pd.DataFrame(lhs(ndim, nsamples - dfInputs.shape[0]), columns=dfInputs.columns)
            x = pd.concat((dfInputs, extra))
    if sampleScheme == "CompFunc":
        x = composition_sampler(ndim, nsamples)
    return np.array(x)


def _given(v):
    return not (isinstance(v, list) and len(v) == 0)


def apply_constraints(samples, ndim, resolution=[], A=[], b=[], Aeq=[], beq=[],
                      lb=[], ub=[], func=[], sampleScheme="LHS", opt_sample_size=True,
                      evaluatedPoints=[]):
    """util.py:104-214: sample the design space and keep the rows whose violation count is zero.

    Semantics kept from the reference, including its conventions: a row is rejected when
    ``A * x <= b`` / ``Aeq * x <= beq`` holds in any column, when any coordinate is below lb or
    above ub, when a constraint function returns non-zero, or when the point has already been
    evaluated by EVERY model in evaluatedPoints.  With opt_sample_size the sample is grown by
    100 * samples per retry until enough rows survive (up to 2000 * ndim times the request)."""
    pending = np.zeros(5)
    pending[0] = 1 if _given(A) else 0
    pending[1] = 1 if _given(Aeq) else 0
    pending[2] = 1 if _given(lb) else 0
    pending[3] = 1 if _given(ub) else 0
    if _given(func):
        pending[4] = len(func) if isinstance(func, list) else 1
    draw = samples
    best_rows, best_count = [], 0
    while True:
        try:
            x = sampleDesignSpace(ndim, draw, sampleScheme)
        except Exception:
            x = sampleDesignSpace(ndim, draw, "LHS")
        if _given(resolution):
            x = np.round(x, decimals=resolution)
        hits = np.zeros((x.shape[0], ndim))
        if _given(A) and _given(b) and len(A) == ndim:
            hits += np.tile(np.array(A), (x.shape[0], 1)) * x <= b
            pending[0] = 0
        if _given(Aeq) and _given(beq):
            hits += np.tile(np.array(Aeq), (x.shape[0], 1)) * x <= beq
            pending[1] = 0
        if _given(lb) and len(lb) == ndim:
            hits += x < np.array(lb).reshape((1, ndim))
            pending[2] = 0
        if _given(ub) and len(ub) == ndim:
            hits += x > np.array(ub).reshape((1, ndim))
            pending[3] = 0
        violations = np.sum(hits, axis=1)
        if isinstance(func, list) and _given(func):
            for fn in func:
                try:
                    violations += fn(x)
                    pending[4] -= 1
                except Exception:
                    pass
        elif not isinstance(func, list):
            try:
                violations += func(x)
                pending[4] = 0
            except Exception:
                pass
        if _given(evaluatedPoints):
            seen = np.zeros_like(violations)
            for pts in evaluatedPoints:
                seen += (x[:, None] == pts).all(-1).any(-1)
            seen[np.where(seen < len(evaluatedPoints))] = 0
            violations += seen
        keep = np.where(violations == 0)[0]
        if not opt_sample_size:
            return x[keep, :], True
        if keep.shape[0] >= samples:
            return x[keep[0:samples], :], bool(np.sum(pending) == 0)
        if len(keep) > best_count:
            best_count, best_rows = len(keep), x[keep, :]
        if draw / samples < ndim * 2000:
            draw += samples * 100
        else:
            return best_rows, False
