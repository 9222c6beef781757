"""The BAREFOOT orchestrator with the reference's public API (barefoot.py:34-2407 of the reference):
``barefoot(...)``, ``initialize_parameters(...)``, ``run_optimization()``.

Host control flow, bookkeeping and output files follow the reference (same constructor keywords,
same ``modelParam`` keys, same ``results/<name>/evaluatedPoints(.csv)`` / ``iterationData(.csv)``
columns, same use of the global numpy RNG), but the three process-pool fan-outs
(barefoot.py:1070-1073 ROM stage, :1271-1283 Truth-Model stage, :2001-2013 batch-only stage)
are ONE device call each over the whole (model, fantasy point, hyper-parameter set) grid
(util.rom_stage_batch, util.fused_calculate_batch, util.batch_acquisition_batch).

Not carried over: the SLURM/file-polling sub-process mechanism
(``multiNode`` is accepted; ranks of a torchrun launch take its role), multi-objective EHVI,
save-state restore (disabled in the reference itself, barefoot.py:311-316) and external
Truth-Model hand-off files."""
import logging
import os
from pickle import dump, load
from time import time

import numpy as np
import pandas as pd

from . import dist as bdist
from .gpModel import gp_model
from .reificationFusion import model_reification
from .util import (apply_constraints, batch_acquisition_batch, call_model, evaluate_fused_model_batch,
                   fused_calculate_batch, kmedoids_max, rom_stage_batch)

_ACQ_NAMES = ["Hedge", "Greedy", "EI", "KG", "TS", "PI", "UCB"]


def _guard(message):
    """barefoot.py:210-224, 1678-1697: log the exception and report success as a bool."""
    def wrap(func):
        def run(self, *args, **kwargs):
            try:
                func(self, *args, **kwargs)
                return True
            except Exception as err:
                self.logger.critical(message)
                self.logger.exception(err)
                return False
        run.__name__ = func.__name__
        run.__doc__ = func.__doc__
        return run
    return wrap


class barefoot():
    def __init__(self, ROMModelList=[], TruthModel=[], calcInitData=True,
                 initDataPathorNum=[], multiNode=0, workingDir=".",
                 calculationName="Calculation", nDim=1, input_resolution=5, restore_calc=False,
                 updateROMafterTM=False, externalTM=False, acquisitionFunc="KG",
                 A=[], b=[], Aeq=[], beq=[], lb=[], ub=[], func=[], keepSubRunning=True,
                 verbose=False, sampleScheme="LHS", tmSampleOpt="Greedy", logname="BAREFOOT",
                 maximize=True, train_func=[], reification=True, batch=True,
                 multiObjective=False, multiObjectRef=[]):
        """Same keywords as the reference constructor (barefoot.py:35-109)."""
        level = logging.DEBUG if verbose else logging.INFO
        # Under torchrun every rank runs this orchestrator (the acquisition grids are sharded inside the
        # batch calls) but the reference has ONE driver process: rank 0 is authoritative.  It alone writes
        # the log and the result files; its host RNG state is adopted by every rank here, and the wall-clock
        # model costs that steer the budgets are broadcast from it (_rank0), so that all ranks draw the same
        # candidates / hyper-parameter sets and take the same branches into the collectives.
        self.rank, self.world = bdist.world()
        self.logger = logging.getLogger(logname if self.rank == 0 else "{}.rank{}".format(logname, self.rank))
        for h in list(self.logger.handlers):
            self.logger.removeHandler(h)
        self.logger.setLevel(level)
        if self.rank == 0:
            handler = logging.FileHandler('{}.log'.format(logname), mode='w')
            handler.setLevel(level)
            handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
        else:
            handler = logging.NullHandler()
            self.logger.propagate = False
        self.logger.addHandler(handler)
        if self.world > 1:
            np.random.set_state(bdist.broadcast_object(np.random.get_state()))
        self.logger.info("Start BAREFOOT Framework Initialization (B200 path) | Calculation Name: {}".format(
            calculationName))
        if multiObjective:
            raise NotImplementedError("multi-objective EHVI is outside the B200 hot path (DESIGN.md section 7)")
        if restore_calc or externalTM:
            raise NotImplementedError("save-state restore / external Truth-Model files are not carried over "
                                      "(the reference's own save is disabled, barefoot.py:311-316)")
        self.restore_calc = False
        self.timeCheck = time()
        self.multiObjective = False
        self.MORef = multiObjectRef
        self.ROM = ROMModelList
        self.TM = TruthModel
        self.TMInitInput, self.TMInitOutput = [], []
        self.ROMInitInput, self.ROMInitOutput = [], []
        self.inputLabels = []
        self.multinode = multiNode
        if multiNode:
            self.logger.warning("multiNode={}: sub-process files are not used; launch with torchrun to shard the "
                                "acquisition grids over GPUs".format(multiNode))
        self.workingDir = workingDir
        self.calculationName = calculationName
        self.calcInitData = calcInitData
        self.initDataPathorNum = initDataPathorNum
        self.currentIteration = -1
        self.maximize = maximize
        if tmSampleOpt in _ACQ_NAMES:
            self.tmSampleOpt = tmSampleOpt
        else:
            self.tmSampleOpt = "Greedy"
            self.logger.warning("Invalid Truth Model Acquisition Function! Using default (Greedy).")
        if acquisitionFunc in _ACQ_NAMES:
            self.acquisitionFunc = acquisitionFunc
        else:
            self.acquisitionFunc = "KG"
            self.logger.warning("Invalid ROM Acquisition Function! Using default (KG).")
        self.nDim = nDim
        self.res = input_resolution
        self.A, self.b, self.Aeq, self.beq, self.ub, self.lb = A, b, Aeq, beq, ub, lb
        self.constr_func = func
        self.train_func = train_func
        if sampleScheme in ["LHS", "Grid", "Custom", "CompFunc"]:
            self.sampleScheme = sampleScheme
        else:
            self.sampleScheme = "LHS"
            self.logger.warning("Invalid Sample Scheme! Using default (LHS).")
        self.keepSubRunning = True
        self.updateROMafterTM = updateROMafterTM
        self.reification = reification
        self.batch = batch
        self.externalTM = False
        self._make_directories()
        self._make_dataframes()
        self.__get_initial_data__()
        self.logger.info("Initialization Completed")

    # ------------------------------------------------------------------------------------------
    # files and dataframes (barefoot.py:226-373)
    # ------------------------------------------------------------------------------------------
    def _make_directories(self):
        if self.rank != 0:
            return
        for sub in ("results", "data", "data/parameterSets", "results/{}".format(self.calculationName)):
            os.makedirs('{}/{}'.format(self.workingDir, sub), exist_ok=True)

    def _make_dataframes(self):
        iter_cols = ["Iteration", "Max Found", "Calculation Time", "Truth Model"]
        iter_cols += ["ROM {}".format(ii) for ii in range(len(self.ROM))]
        self.inputLabels = ["x{}".format(ii) for ii in range(self.nDim)]
        self.evaluatedPoints = pd.DataFrame(columns=["Model Index", "Iteration", "y"] + self.inputLabels)
        self.iterationData = pd.DataFrame(columns=iter_cols)

    @staticmethod
    def _rank0(value):
        """Rank 0's value of a host quantity that steers the control flow (wall-clock costs)."""
        return bdist.broadcast_object(value)

    def _model_result(self, job):
        """call_model (util.py:51-55).  Under torchrun only rank 0 evaluates the user's model (it may be
        expensive, stateful or not deterministic) and every rank receives its result."""
        if self.world == 1:
            return call_model(job)
        out = None
        if self.rank == 0:
            try:
                out = call_model(job)
            except Exception as err:                           # noqa: BLE001 - re-raised on every rank
                out = err
        out = bdist.broadcast_object(out)
        if isinstance(out, Exception):
            raise out
        return out

    def _save_dataframes(self):
        if self.rank != 0:
            return
        base = '{}/results/{}'.format(self.workingDir, self.calculationName)
        for name, frame in (("evaluatedPoints", self.evaluatedPoints), ("iterationData", self.iterationData)):
            with open('{}/{}'.format(base, name), 'wb') as f:
                dump(frame, f)
            frame.to_csv('{}/{}.csv'.format(base, name))
        if self.acquisitionFunc == "Hedge" or self.tmSampleOpt == "Hedge":
            record = {"ROM": self.gpHedgeTrack if self.acquisitionFunc == "Hedge" else [],
                      "TM": self.gpHedgeTrackTM if self.tmSampleOpt == "Hedge" else []}
            with open('{}/hedgeRecord'.format(base), 'wb') as f:
                dump(record, f)
        self.logger.info("Dataframes Pickled and Dumped to Results Directory")

    def _add_points(self, modelIndex, eval_x, eval_y):
        rows = np.zeros((eval_x.shape[0], self.nDim + 3))
        rows[:, 0] = modelIndex
        rows[:, 1] = self.currentIteration
        rows[:, 2] = eval_y
        rows[:, 3:] = eval_x
        rows = pd.DataFrame(rows, columns=self.evaluatedPoints.columns)
        self.evaluatedPoints = rows if self.evaluatedPoints.shape[0] == 0 else pd.concat([self.evaluatedPoints, rows])

    def _add_iteration(self, calcTime, counts):
        row = np.zeros((1, 4 + len(self.ROM)))
        row[0, 0] = self.currentIteration
        row[0, 1] = self.maxTM
        row[0, 2] = calcTime
        row[0, 3] = counts[-1]
        row[0, 4:] = counts[0:len(self.ROM)]
        row = pd.DataFrame(row, columns=self.iterationData.columns)
        self.iterationData = row if self.iterationData.shape[0] == 0 else pd.concat([self.iterationData, row])

    def _last_counts(self):
        """Model-call counts of the latest iteration row, as [ROM 0.., Truth Model] (barefoot.py:1313-1318)."""
        current = np.array(self.iterationData.iloc[:, 3:])[-1, :]
        count = np.zeros((len(self.ROM) + 1))
        count[0:len(self.ROM)] = current[1:]
        count[-1] = current[0]
        return count

    def _sample(self, count, scheme=None, opt_sample_size=True, evaluated=[]):
        return apply_constraints(count, self.nDim, resolution=self.res, A=self.A, b=self.b, Aeq=self.Aeq,
                                 beq=self.beq, lb=self.lb, ub=self.ub, func=self.constr_func,
                                 sampleScheme=self.sampleScheme if scheme is None else scheme,
                                 opt_sample_size=opt_sample_size, evaluatedPoints=evaluated)

    def _evaluated_inputs(self, model_indices):
        return [np.array(self.evaluatedPoints.loc[self.evaluatedPoints['Model Index'] == pp, self.inputLabels])
                for pp in model_indices]

    # ------------------------------------------------------------------------------------------
    # initial data (barefoot.py:375-586)
    # ------------------------------------------------------------------------------------------
    @_guard("Initialization Code Failed - See Error Below")
    def __get_initial_data__(self):
        self.maxTM = -np.inf
        if self.acquisitionFunc == "Hedge":
            self.gpHedgeHist = [[np.random.random()] for _ in range(6)]
            self.gpHedgeProb = np.sum(self.gpHedgeHist, axis=1)
            self.gpHedgeTrack = []
        if self.tmSampleOpt == "Hedge":
            self.gpHedgeHistTM = [[np.random.random()] for _ in range(6)]
            self.gpHedgeProbTM = np.sum(self.gpHedgeHistTM, axis=1)
            self.gpHedgeTrackTM = []
        self.goal = 1 if self.maximize else -1
        if self.calcInitData:
            jobs = []
            if self.reification:
                for ii in range(len(self.ROM)):
                    initInput, ok = self._sample(self.initDataPathorNum[ii])
                    if not ok:
                        self.logger.critical("ROM {} - Initial Data - constraints not met, continuing with {}/{}".format(
                            ii, initInput.shape[0], self.initDataPathorNum[ii]))
                    jobs += [{"Model Index": ii, "Model": self.ROM[ii], "Input Values": initInput[jj, :]}
                             for jj in range(initInput.shape[0])]
                    self.ROMInitInput.append(np.zeros_like(initInput))
                    self.ROMInitOutput.append(np.zeros(initInput.shape[0]))
            initInput, ok = self._sample(self.initDataPathorNum[-1])
            if not ok:
                self.logger.critical("TM - Initial Data - constraints not met, continuing with {}/{}".format(
                    initInput.shape[0], self.initDataPathorNum[-1]))
            jobs += [{"Model Index": -1, "Model": self.TM, "Input Values": initInput[jj, :]}
                     for jj in range(initInput.shape[0])]
            self.TMInitInput = np.zeros_like(initInput)
            self.TMInitOutput = np.zeros(initInput.shape[0])
            filled = [0] * (len(self.ROM) + 1) if self.reification else [0]
            xs, ys, idx = [], [], []
            for job in jobs:
                results = self._model_result(job)
                if not hasattr(results, "shape"):
                    continue                       # a failed model call returns a non-array and is dropped
                m = job["Model Index"]
                if m != -1:
                    self.ROMInitInput[m][filled[m], :] = job["Input Values"]
                    self.ROMInitOutput[m][filled[m]] = np.asarray(self.goal * results).reshape(-1)[0]
                else:
                    self.TMInitInput[filled[m], :] = job["Input Values"]
                    self.TMInitOutput[filled[m]] = np.asarray(self.goal * results).reshape(-1)[0]
                    if np.max(results) > self.maxTM:
                        self.maxTM = np.max(results)
                filled[m] += 1
                xs.append(job["Input Values"])
                ys.append(np.asarray(self.goal * results).reshape(-1)[0])
                idx.append(m)
            temp_x, temp_y, temp_index = np.array(xs).reshape(-1, self.nDim), np.array(ys), np.array(idx, dtype=float)
            counts = np.array(filled)
        else:
            with open(self.initDataPathorNum, 'rb') as f:
                data = load(f)
            self.TMInitOutput, self.TMInitInput = data["TMInitOutput"], data["TMInitInput"]
            if self.reification:
                self.ROMInitOutput, self.ROMInitInput = data["ROMInitOutput"], data["ROMInitInput"]
            xs, ys, idx, counts = [], [], [], []
            if self.reification:
                for ii in range(len(self.ROM)):
                    for jj in range(self.ROMInitOutput[ii].shape[0]):
                        xs.append(self.ROMInitInput[ii][jj, :])
                        ys.append(self.goal * self.ROMInitOutput[ii][jj])
                        idx.append(ii)
                    counts.append(self.ROMInitInput[ii].shape[0])
            for jj in range(self.TMInitOutput.shape[0]):
                xs.append(self.TMInitInput[jj, :])
                ys.append(self.TMInitOutput[jj])
                idx.append(-1)
                if self.TMInitOutput[jj] > self.maxTM:
                    self.maxTM = self.TMInitOutput[jj]
            counts.append(self.TMInitInput.shape[0])
            temp_x, temp_y, temp_index = np.array(xs).reshape(-1, self.nDim), np.array(ys), np.array(idx, dtype=float)
            counts = np.array(counts)
        self._add_points(temp_index, temp_x, temp_y)
        self._add_iteration(time() - self.timeCheck, counts)
        self.timeCheck = time()

    # ------------------------------------------------------------------------------------------
    # parameters (barefoot.py:589-780)
    # ------------------------------------------------------------------------------------------
    @_guard("Initialization Code Failed - See Error Below")
    def initialize_parameters(self, modelParam, covFunc="M32", iterLimit=100,
                              sampleCount=50, hpCount=100, batchSize=5,
                              tmIter=1e6, totalBudget=1e16, tmBudget=1e16,
                              upperBound=1, lowBound=0.0001, fusedPoints=500,
                              fusedHP=[], fusedSamples=10000):
        """Same keywords and meaning as the reference (barefoot.py:589-640).  modelParam keys:
        model_l, model_sf, model_sn, means, std, err_l, err_sf, err_sn, costs."""
        self.covFunc, self.iterLimit, self.sampleCount = covFunc, iterLimit, sampleCount
        self.hpCount, self.batchSize, self.tmIterLim = hpCount, batchSize, tmIter
        self.totalBudget, self.tmBudget = totalBudget, tmBudget
        self.upperBound, self.lowBound = upperBound, lowBound
        self.modelParam = modelParam
        self.modelCosts = modelParam["costs"]
        self.fusedHP, self.fusedSamples = fusedHP, fusedSamples
        # log-decade grid of candidate hyper-parameter values (hpCount per decade), then hpCount sets
        # of (sigma_f, l_1 .. l_d), every entry an independent np.random.randint draw (quirk Q15)
        top = self.lowBound * 10
        grid = np.linspace(self.lowBound, top, num=self.hpCount)
        while top < self.upperBound:
            bottom, top = top, min(top * 10, self.upperBound)
            grid = np.append(grid, np.linspace(bottom, top, num=self.hpCount))
        self.fusedModelHP = np.zeros((self.hpCount, self.nDim + 1))
        for i in range(self.hpCount):
            for j in range(self.nDim + 1):
                self.fusedModelHP[i, j] = grid[np.random.randint(0, grid.shape[0])]
        self.xFused, ok = self._sample(fusedPoints, scheme="CompFunc" if self.sampleScheme == "CompFunc" else "LHS",
                                       opt_sample_size=False)
        if self.reification:
            p = self.modelParam
            self.reificationObj = model_reification(self.ROMInitInput, self.ROMInitOutput, p['model_l'],
                                                    p['model_sf'], p['model_sn'], p['means'], p['std'],
                                                    p['err_l'], p['err_sf'], p['err_sn'], self.TMInitInput,
                                                    self.TMInitOutput, len(self.ROM), self.nDim, self.covFunc)
        else:
            self.modelGP = gp_model(self.TMInitInput, self.TMInitOutput, np.ones((self.nDim)), 1, 0.05,
                                    self.nDim, self.covFunc)
        self.allTMInput, self.allTMOutput = [], []
        self.tmBudgetLeft, self.totalBudgetLeft = self.tmBudget, self.totalBudget
        self.currentIteration += 1
        self.tmIterCount = 0
        self.logger.info("Reification Object Initialized. Ready for Calculations")

    # ------------------------------------------------------------------------------------------
    # ROM stage (barefoot.py:985-1075, 2102-2172, 1308-1396)
    # ------------------------------------------------------------------------------------------
    def _rom_stage(self, x_test, new_mean):
        """All (model, fantasy point, hyper-parameter set) work items in one device pass; rows in the
        reference's loop order.  For Hedge: the six row lists of the portfolio."""
        hp = self.fusedModelHP if self.batch else np.asarray(self.fusedHP, dtype=np.float64)[None]
        out = rom_stage_batch(self.reificationObj, self.xFused, hp, self.covFunc, x_test, new_mean,
                              self.acquisitionFunc, self.modelParam['costs'], self.maxTM,
                              iteration=self.currentIteration + 1, true_sample_count=self.sampleCount,
                              kk_list=list(range(self.sampleCount)), as_arrays=True)
        return out

    @staticmethod
    def _cluster_rows_host(kg_output, n_models):
        """The reference's dictionary loop, verbatim semantics (barefoot.py:2110-2140); kept as the
        statement of what `_cluster_rows` computes on the device and used by the tests."""
        best = {}
        for iii in range(kg_output.shape[0]):
            key, model = kg_output[iii, 2], int(kg_output[iii, 3])
            entry = best.setdefault(key, {'models': [], 'nu': [1e-6] * n_models, 'row': [-1] * n_models})
            if model in entry['models']:
                if kg_output[iii, 1] > entry['nu'][model]:
                    entry['nu'][model], entry['row'][model] = kg_output[iii, 1], iii
            else:
                entry['models'].append(model)
                entry['nu'][model], entry['row'][model] = kg_output[iii, 1], iii
        rows = [[entry['nu'][m], key, m, entry['row'][m]] for key, entry in best.items() for m in entry['models']]
        return np.array(rows, dtype=np.float64)

    @staticmethod
    def _cluster_rows(kg_output, n_models):
        """barefoot.py:2110-2140 on the device (engine.group_best): rows [value, x*, model, row id] for
        every (x*, model) pair that occurs, in the reference's order (x* by first appearance, models
        by first appearance within x*).  Only the final ordering of the <= m N surviving rows is host work."""
        from . import engine
        values = np.asarray(kg_output[:, 1], dtype=np.float64)
        xstar = np.asarray(kg_output[:, 2], dtype=np.float64)
        model = np.asarray(kg_output[:, 3], dtype=np.int64)
        xi = xstar.astype(np.int64)
        if xi.size and (not np.array_equal(xi, xstar) or xi.min() < 0 or model.min() < 0 or model.max() >= n_models):
            raise ValueError("clustering records need integer candidate indices and model indices < %d" % n_models)
        n_x = int(xi.max()) + 1 if xi.size else 1
        first, best_val, best_row = engine.group_best(xi * n_models + model, values, n_x * n_models)
        first, best_val, best_row = first.cpu().numpy(), best_val.cpu().numpy(), best_row.cpu().numpy()
        present = np.where(first < values.shape[0])[0]
        key_first = first.reshape(n_x, n_models).min(axis=1)              # first appearance of x*
        order = present[np.lexsort((first[present], key_first[present // n_models]))]
        return np.column_stack([best_val[order], (order // n_models).astype(np.float64),
                                (order % n_models).astype(np.float64), best_row[order].astype(np.float64)])

    def _cluster_rom(self, kg_output):
        """barefoot.py:2102-2172: per selected candidate x* keep, for every model that proposed it,
        the row with the largest value; cluster the (value, x*, model) rows and return the original
        rows of the chosen ones."""
        med_input = self._cluster_rows(kg_output, len(self.ROM))
        if self.batch:
            k = self.batchSize if med_input.shape[0] > self.batchSize else med_input.shape[0]
            medoids = kmedoids_max(med_input[:, 0:3], k)
        else:
            medoids = [np.where(med_input[:, 0] == np.max(med_input[:, 0]))[0][0]]
        return kg_output[[int(med_input[m, 3]) for m in medoids], :]

    def _call_models(self, jobs, count, on_result):
        """Evaluate model calls in order; failed calls (non-array results) are dropped."""
        xs, ys, idx, passed = [], [], [], 0
        costs = np.zeros(len(jobs))
        for n, job in enumerate(jobs):
            results = self._model_result(job)
            costs[n] += self.modelCosts[job["Model Index"]]
            if not hasattr(results, "shape"):
                continue
            passed += 1
            if len(results.shape) == 1:
                results = np.expand_dims(results, axis=0)
            results = self.goal * results
            xs.append(job["Input Values"])
            ys.append(np.asarray(results).reshape(-1)[0])
            idx.append(job["Model Index"])
            on_result(job, results)
            count[job["Model Index"]] += 1
        return (np.array(xs).reshape(-1, self.nDim), np.array(ys), np.array(idx, dtype=float), costs, count, passed)

    def _call_ROM(self, medoid_out, x_val):
        jobs = [{"Model Index": int(medoid_out[iii, 3]), "Model": self.ROM[int(medoid_out[iii, 3])],
                 "Input Values": x_val[iii, :]} for iii in range(medoid_out.shape[0])]
        self.logger.info("Start ROM Function Evaluations | {} Calculations".format(len(jobs)))
        return self._call_models(jobs, self._last_counts(),
                                 lambda job, res: self.reificationObj.update_GP(job["Input Values"], res,
                                                                                job["Model Index"]))

    def _call_Truth(self, jobs, count):
        def update(job, res):
            if self.reification:
                self.reificationObj.update_truth(job["Input Values"], res)
            else:
                self.modelGP.update(job["Input Values"], res, 0.05, False)
        self.logger.info("Start Truth Model Evaluations | {} Sets".format(len(jobs)))
        temp_x, temp_y, temp_index, costs, count, passed = self._call_models(jobs, count, update)
        if passed:
            self._add_points(temp_index, temp_x, temp_y)
            self.totalBudgetLeft -= self.batchSize * self.modelCosts[-1]
            if np.max(temp_y) > self.maxTM:
                self.maxTM = np.max(temp_y)
        else:
            self.logger.critical("All Truth Model Evaluations Failed to Produce Results! Continue with no new results.")
        return count

    # ------------------------------------------------------------------------------------------
    # Truth-Model stage (barefoot.py:1201-1306, 1618-1676)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _dedupe(pairs, n_candidates):
        """barefoot.py:1290-1300: best value per selected index; value 0 means empty (quirk Q13)."""
        max_values = np.zeros((n_candidates, 2))
        for value, index in pairs:
            index = int(index)
            if max_values[index, 0] != 0:
                if max_values[index, 0] < value:
                    max_values[index, 0], max_values[index, 1] = value, index
            else:
                max_values[index, 0], max_values[index, 1] = value, index
        out = max_values[np.where(max_values[:, 0] != 0)]
        return out if out.shape[0] else np.array([[0, 0]])

    def _tm_stage(self, tm_test):
        hp = self.fusedModelHP if self.batch else np.asarray(self.fusedHP, dtype=np.float64)[None]
        out = fused_calculate_batch(self.reificationObj, self.xFused, hp, self.covFunc, tm_test, self.maxTM,
                                    self.tmSampleOpt, iteration=self.currentIteration + 1, xi=0.01)
        if self.tmSampleOpt != "Hedge":
            return self._dedupe(zip(out[0], out[1]), tm_test.shape[0])
        # GP-Hedge: choose the portfolio member with the largest gain, cluster all six for the gain update
        portfolio = [[item[k] for item in out] for k in range(6)]
        prob = self.gpHedgeProbTM / np.sum(self.gpHedgeProbTM)
        chosen = np.where(prob == np.max(prob))[0][0]
        self.gpHedgeTrackTM.append(prob)
        fused_output = self._dedupe(portfolio[chosen], tm_test.shape[0])
        clusters = []
        for k in range(6):
            cluster_output = np.array(portfolio[k], dtype=object)
            try:
                if cluster_output.shape[0] > self.batchSize:
                    medoids = kmedoids_max(cluster_output, self.batchSize)
                else:
                    medoids = list(range(cluster_output.shape[0]))
            except Exception:
                medoids = kmedoids_max(cluster_output, 1)
            clusters.append(np.array(tm_test[medoids, :], dtype=float))
        self._hedge_clusters = clusters
        return fused_output

    def _update_hedge(self, which):
        """barefoot.py:1527-1601: gain of each portfolio member = max over its proposed points of the
        fused (or single-GP) mean averaged over the hyper-parameter sets."""
        model = self.reificationObj if self.reification else self.modelGP
        gains = []
        for k in range(6):
            mu = evaluate_fused_model_batch(model, self.reification, self.xFused, self.fusedModelHP, self.covFunc,
                                            np.array(self._hedge_clusters[k]))
            gains.append(np.max(np.mean(mu.transpose(), axis=1)))
        hist = self.gpHedgeHist if which == "ROM" else self.gpHedgeHistTM
        for k in range(6):
            hist[k].append(gains[k])
            if len(hist[k]) > 2 * self.tmIterLim:
                hist[k] = hist[k][1:]
        if which == "ROM":
            self.gpHedgeProb = np.sum(self.gpHedgeHist, axis=1)
        else:
            self.gpHedgeProbTM = np.sum(self.gpHedgeHistTM, axis=1)

    # ------------------------------------------------------------------------------------------
    # main loops
    # ------------------------------------------------------------------------------------------
    def _finished(self):
        return (self.totalBudgetLeft < 0) or (self.currentIteration >= self.iterLimit)

    @_guard("Optimization Code Failed - See Error Below")
    def run_BAREFOOT(self):
        """barefoot.py:1700-1918: ROM stage -> clustering -> ROM calls -> (when due) Truth-Model stage."""
        self.logger.info("Start Full BAREFOOT Framework Calculation" if self.batch
                         else "Start Reification Only Framework Calculation")
        while True:
            self.logger.info("Start Iteration : {:04d}".format(self.currentIteration))
            self.timeCheck = time()
            x_test, ok = self._sample(self.sampleCount, evaluated=self._evaluated_inputs(range(len(self.ROM))))
            if not ok:
                self.logger.critical("ROM - Sample Size NOT met due to constraints! Continue with {}/{} Samples".format(
                    x_test.shape[0], self.sampleCount))
            new_mean = [self.reificationObj.predict_low_order(x_test, iii)[0] for iii in range(len(self.ROM))]
            out = self._rom_stage(x_test, new_mean)
            if self.acquisitionFunc == "Hedge":
                prob = self.gpHedgeProb / np.sum(self.gpHedgeProb)
                chosen = np.where(prob == np.max(prob))[0][0]
                self.gpHedgeTrack.append(chosen)
                # The reference's pool returns one entry per WORK ITEM (each holding the six portfolio
                # rows) and then indexes that list with the portfolio index (barefoot.py:1509-1521): the
                # rows of work item `chosen` are clustered, and work items 0..5 feed the gain update.
                # Reproduced as is.
                items = [np.stack([out[k][s_] for k in range(6)]) for s_ in range(6)]
                self._hedge_clusters = [x_test[self._cluster_rom(items[k])[:, 2].astype(int), :] for k in range(6)]
                kg_output = items[chosen]
            else:
                kg_output = out
            medoid_out = self._cluster_rom(kg_output)
            model_cost = self._rank0(time() - self.timeCheck)
            self.timeCheck = time()
            temp_x, temp_y, temp_index, costs, count, passed = self._call_ROM(
                medoid_out, x_test[medoid_out[:, 2].astype(int), :])
            if passed != 0:
                self._add_points(temp_index, temp_x, temp_y)
                if self.acquisitionFunc == "Hedge":
                    self._update_hedge("ROM")
            else:
                self.logger.critical("All ROM Evalutions Failed to produce a result! Continue with no new data")
            self.totalBudgetLeft -= np.sum(costs) + model_cost
            self.tmBudgetLeft -= np.sum(costs) + model_cost
            tm_due = (self.tmBudgetLeft < 0) or (self.tmIterCount == self.tmIterLim)
            if tm_due:
                self.logger.info("Start Truth Model Evaluations")
                tm_test, ok = self._sample(self.fusedSamples, evaluated=self._evaluated_inputs([-1]))
                fused_output = np.array(self._tm_stage(tm_test))
                if self.batch:
                    if fused_output.shape[0] > self.batchSize:
                        medoids = kmedoids_max(fused_output, self.batchSize)
                    else:
                        medoids = list(range(fused_output.shape[0])) if self.batchSize != 0 else []
                else:
                    medoids = [np.where(fused_output[:, 0] == np.max(fused_output[:, 0]))[0][0]]
                jobs = [{"Model Index": -1, "Model": self.TM,
                         "Input Values": np.array(tm_test[int(fused_output[m, 1]), :], dtype=float)} for m in medoids]
                for _ in range(self.batchSize - len(medoids)):      # top up with random candidates
                    jobs.append({"Model Index": -1, "Model": self.TM,
                                 "Input Values": np.array(tm_test[np.random.randint(0, tm_test.shape[0]), :],
                                                          dtype=float)})
                self.tmIterCount = 0
                self.tmBudgetLeft = self.tmBudget
                count = self._call_Truth(jobs, count)
                if self.tmSampleOpt == "Hedge":
                    self._update_hedge("TM")
            self._add_iteration(time() - self.timeCheck + model_cost, count)
            self.timeCheck = time()
            self._save_dataframes()
            if (self.tmBudgetLeft < 0) or (self.tmIterCount == self.tmIterLim):
                if self.updateROMafterTM:
                    self._update_reduced_order_models()
            self.logger.info("Iteration {} Completed Successfully".format(self.currentIteration))
            done = self._finished()
            self.currentIteration += 1
            self.tmIterCount += 1
            if done:
                self.logger.info("Iteration or Budget Limit Met or Exceeded | BAREFOOT Calculation Completed")
                break

    @_guard("Optimization Code Failed - See Error Below")
    def run_BATCH(self):
        """barefoot.py:1920-2092: batch Bayesian optimisation on the Truth Model alone."""
        self.logger.info("Start Batch Only Framework Calculation")
        while True:
            self.logger.info("Start Iteration : {:04d}".format(self.currentIteration))
            self.timeCheck = time()
            x_test, ok = self._sample(self.sampleCount, evaluated=self._evaluated_inputs(range(len(self.ROM))))
            count = self._last_counts()
            out = batch_acquisition_batch(self.modelGP, self.fusedModelHP, x_test, self.maxTM, self.tmSampleOpt,
                                          iteration=self.currentIteration + 1)

            def get_medoids(rows):
                if rows.shape[0] > self.batchSize:
                    return kmedoids_max(rows[:, 0:3], self.batchSize)
                return list(range(rows.shape[0]))

            if self.tmSampleOpt == "Hedge":
                prob = self.gpHedgeProbTM / np.sum(self.gpHedgeProbTM)
                chosen = np.where(prob == np.max(prob))[0][0]
                self.gpHedgeTrackTM.append(prob)
                clusters = []
                for k in range(6):
                    rows = np.unique(np.array([item[k] for item in out]), axis=0)
                    med = get_medoids(rows)
                    if k == chosen:
                        medoids, kg_output = med, rows
                    index = np.array(rows[med, 1], dtype=np.uint8)          # uint8 as in the reference (:2034)
                    clusters.append(np.array(x_test[index, :], dtype=float))
                self._hedge_clusters = clusters
            else:
                kg_output = np.unique(np.column_stack([out[0], np.asarray(out[1], dtype=np.float64)]), axis=0)
                medoids = get_medoids(kg_output)
            model_cost = self._rank0(time() - self.timeCheck)
            self.timeCheck = time()
            jobs = [{"Model Index": -1, "Model": self.TM,
                     "Input Values": np.array(x_test[int(kg_output[m, 1]), :], dtype=float)} for m in medoids]
            count = self._call_Truth(jobs, count)
            if self.acquisitionFunc == "Hedge":
                self._update_hedge("TM")
            self._add_iteration(time() - self.timeCheck + model_cost, count)
            self.timeCheck = time()
            self._save_dataframes()
            self.logger.info("Iteration {} Completed Successfully".format(self.currentIteration))
            done = self._finished()
            self.currentIteration += 1
            self.tmIterCount += 1
            if done:
                self.logger.info("Iteration or Budget Limit Met or Exceeded | BAREFOOT Calculation Completed")
                break

    def run_optimization(self):
        """barefoot.py:2094-2099."""
        if self.reification:
            return self.run_BAREFOOT()
        return self.run_BATCH()

    def _update_reduced_order_models(self):
        """barefoot.py:2248-2336 (retrain user ROMs through train_func, then rebuild the ROM GPs)."""
        raise NotImplementedError("updateROMafterTM is outside the B200 hot path (DESIGN.md section 7)")
