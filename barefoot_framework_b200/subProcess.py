"""Worker for the reference's multi-node file protocol (subProcess.py:17-191 of the reference), with
the work items evaluated on this process's GPU.

    python -m barefoot_framework_b200.subProcess <index> [poll-seconds]

The reference master (barefoot.py:808-983, 1077-1199) writes ``subprocess/<i>.dump`` (a list of
[parameterIndex, parameterFileIndex] work items), sets ``subprocess/sub<i>.control`` to
[0, "iteration" | "fused", acquisition] and polls for ``subprocess/<i>.output``; this worker keeps that
contract so it can serve an unmodified master.  Inside one torchrun job the same role is played
without files by barefoot_framework_b200.dist (set / candidate sharding + an all-gather of the
per-set records), which is what the in-tree orchestrator uses."""
import logging
import os
import sys
from pickle import dump, load
from time import sleep, time

import numpy as np

from . import util

_ROM_FUNCS = {"KG": util.calculate_KG, "EI": util.calculate_EI, "TS": util.calculate_TS,
              "Hedge": util.calculate_GPHedge, "Greedy": util.calculate_Greedy, "PI": util.calculate_PI,
              "UCB": util.calculate_UCB}


def _dedupe(fused_output, n_candidates):
    """subProcess.py:149-160: best value per selected index, value 0 means empty."""
    max_values = np.zeros((n_candidates, 2))
    for value, index in fused_output:
        index = int(index)
        if max_values[index, 0] != 0:
            if max_values[index, 0] < value:
                max_values[index, 0], max_values[index, 1] = value, index
        else:
            max_values[index, 0], max_values[index, 1] = value, index
    return max_values[np.where(max_values[:, 0] != 0)]


def serve_once(index, logger=None):
    """Handle one pending calculation of worker `index`; returns True if one was run."""
    control_path = "subprocess/sub{}.control".format(index)
    if not os.path.isfile(control_path):
        return False
    with open(control_path, 'rb') as f:
        control_param = load(f)
    if control_param[0] != 0:
        return False
    start = time()
    with open("subprocess/{}.dump".format(index), 'rb') as f:
        parameters = load(f)
    kind, acq = control_param[1], control_param[2]
    if acq == "EHVI":
        raise NotImplementedError("multi-objective EHVI is outside the B200 hot path")
    if kind == "iteration":
        output = [_ROM_FUNCS[acq](p) for p in parameters]
    elif kind == "fused":
        results = [util.fused_calculate(p) for p in parameters]
        if acq == "Hedge":
            # the reference worker keeps the first four portfolio members only (subProcess.py:113,131-134)
            output = [[r[0][k] for r in results] for k in range(4)]
        else:
            output = _dedupe([r[0] for r in results], results[-1][1])
    else:
        raise ValueError("unknown calculation type {!r}".format(kind))
    with open("subprocess/{}.output".format(index), 'wb') as f:
        dump(output, f)
    control_param[0] = 1
    with open(control_path, 'wb') as f:
        dump(control_param, f)
    if logger:
        logger.info("{} | Calculation Results Dumped | {} hours".format(index, np.round((time() - start) / 3600, 4)))
    return True


def main(argv=None):
    argv = sys.argv if argv is None else argv
    index = argv[1]
    poll = float(argv[2]) if len(argv) > 2 else 10.0
    logger = logging.getLogger('BAREFOOT.subprocess')
    logger.setLevel(logging.INFO)
    handler = logging.FileHandler('BAREFOOT.log')
    handler.setFormatter(logging.Formatter('%(asctime)s - %(name)s - %(levelname)s - %(message)s'))
    logger.addHandler(handler)
    logger.info("Subprocess {} | started".format(index))
    with open("subprocess/sub{}.start".format(index), 'w') as f:
        f.write("subprocess started successfully\n\n")
    while not os.path.isfile('subprocess/close{}'.format(index)):
        try:
            serve_once(index, logger)
        except Exception as exc:            # the reference worker logs and keeps polling (subProcess.py:172-175)
            logger.critical("Error completing Calculation | {}".format(exc))
            logger.exception(exc)
        sleep(poll)
    with open("subprocess/sub{}.start".format(index), 'a') as f:
        f.write("subprocess finished successfully\n\n")
    logger.info("{} | Subprocess Finished".format(index))


if __name__ == "__main__":
    main()
