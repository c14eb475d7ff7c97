#!/usr/bin/env python3
"""B200-native drop-in for ``ppanggolin/nem/partition.py`` (PPanGGOLiN 2.3.1).

Same public names, arguments, return conventions and side effects as the reference module
(/root/reference/ppanggolin/nem/partition.py): ``partition``, ``evaluate_nb_partitions``, ``partition_nem``,
``run_partitioning``, ``write_nem_input_files``, ``nem_single``, ``nem_samples``, ``launch``, ``subparser``,
``parser_partition`` and the module globals ``pan`` / ``samples``.  What changes is below the API:

* ``write_nem_input_files`` exports in memory (bit-packed matrix + CSR, ``export.py``) instead of writing text files;
  the "directory" argument is kept as the handle that ``run_partitioning`` receives (text files are only written with
  ``keep_tmp_files``).
* ``run_partitioning`` calls the sm_100a CUDA path through the C ABI (``engine.py``) instead of ``nem_stats.nem`` and
  rebuilds what the reference parses out of ``.uf``/``.mf`` (3-decimal posteriors, ``%g`` criteria) on the device.
* no fork pools: a CUDA context does not survive ``fork``; independent K / chunk instances run back to back on the
  GPU (``cpu`` is accepted and ignored).

There is no CPU fallback: without the CUDA library the functions raise.
"""
from __future__ import annotations

import argparse
import logging
import math
import os
import random
import tempfile
import time
from collections import Counter, defaultdict
from pathlib import Path
from shutil import copytree
from typing import List, Tuple, Union

import numpy as np

from . import engine, export

try:  # inside an installed PPanGGOLiN these resolve to the real data model / persistence layer
    from ppanggolin.pangenome import Pangenome
    from ppanggolin.utils import mk_outdir
    from ppanggolin.formats import check_pangenome_info, write_pangenome, erase_pangenome
except Exception:  # standalone use (tests, benchmarks): the callers hand over in-memory objects
    class Pangenome:  # minimal stand-in for the module-level default object (partition.py:31)
        def __init__(self):
            self.status = defaultdict(lambda: "No")
            self.parameters = {}

    def mk_outdir(output, force: bool = False, exist_ok: bool = False):
        Path(output).mkdir(parents=True, exist_ok=True)

    def check_pangenome_info(pangenome, **kwargs):
        return None

    def write_pangenome(*args, **kwargs):
        raise RuntimeError("write_pangenome needs the PPanGGOLiN persistence layer (ppanggolin.formats)")

    def erase_pangenome(*args, **kwargs):
        raise RuntimeError("erase_pangenome needs the PPanGGOLiN persistence layer (ppanggolin.formats)")

pan = Pangenome()
samples = []

# D above which the reference's double-precision exp() underflows for every class; beyond it
# the log-sum-exp epilogue is used instead of the reference-arithmetic one.
REF_EPILOGUE_MAX_ORGS = 1500


def _epilogue_flags(nb_org: int) -> int:
    forced = os.environ.get("PPANGGOLIN_B200_EPILOGUE", "")
    if forced == "ref":
        return engine.EPI_REF
    if forced == "logspace":
        return engine.EPI_LOGSPACE
    return engine.EPI_REF if nb_org <= REF_EPILOGUE_MAX_ORGS else engine.EPI_LOGSPACE


def _device_input(inp):
    if inp.bits_tensor is not None and inp.bits_tensor.is_cuda:
        return engine.DeviceInput.from_tensor(inp.bits_tensor, inp.D, inp.rowptr, inp.col, inp.w)
    return engine.DeviceInput.from_numpy(inp.host_bits(), inp.D, inp.rowptr, inp.col, inp.w)


# One instance whose matrix is at least this large is row-sharded over the ranks of an initialised process group (SURVEY.md
# section 8e, "one large instance"); smaller ones run on the calling rank's GPU.
SHARD_MIN_BYTES = int(os.environ.get("PPANGGOLIN_B200_SHARD_MIN_BYTES", str(256 << 20)))


def _run_sharded(dev_in, D: int, names, kval: int, beta: float, free_dispersion: bool, itermax: int, flags: int):
    """The instance through ``nem/sharded.py`` when a process group is active and the matrix is large: every rank keeps a
    contiguous block of families, the exchange runs over NVLink peer memory inside the kernels, and every rank ends with
    the same posteriors as a single-GPU run.  Returns None when the instance should run on one GPU."""
    rank, world, dist = _rank_world()
    N, Ws = dev_in.bits.shape
    if world == 1 or N * Ws * 4 < SHARD_MIN_BYTES or N < 2 * world:
        return None
    # every rank must hold the same export (same family rows, same genome columns): column order comes from the iteration
    # order of the caller's genome collection, which may differ between processes -- compare a digest first and let every
    # rank run the whole instance itself if they differ (same answer, no speed-up)
    import hashlib
    import torch
    colsum = torch.zeros(Ws, dtype=torch.int64, device=dev_in.bits.device)     # per 32-column word: sum of the rows' words
    for r0 in range(0, N, 8192):
        colsum += (dev_in.bits[r0:r0 + 8192].to(torch.int64) & 0xFFFFFFFF).sum(0)
    tail = "|".join(list(names[:64]) + list(names[-64:])) if names is not None else ""
    digest = hashlib.sha256(colsum.cpu().numpy().tobytes() + tail.encode()).hexdigest()
    seen = [None] * world
    dist.all_gather_object(seen, digest)
    if len(set(seen)) != 1:
        logging.getLogger("PPanGGOLiN").warning("The ranks hold differently ordered exports of this NEM instance; "
                                                "it runs unsharded on every rank.")
        return None
    from .sharded import CudaBackend, ShardedEM, shard_rows
    beta32 = float(np.float32(beta))
    lo, hi = shard_rows(N, world, rank)
    graph = dev_in if beta32 != 0.0 else None
    # the peer-memory exchange maps the other ranks' buffers through CUDA IPC: one host, at most engine.MAX_PEERS ranks;
    # anything else exchanges through the NCCL collectives of ShardedEM (agreed on by all ranks: hostnames are gathered)
    import socket
    hosts = [None] * world
    dist.all_gather_object(hosts, socket.gethostname())
    peer_ok = world <= engine.MAX_PEERS and len(set(hosts)) == 1
    be = CudaBackend(dev_in.bits[lo:hi], D, N, kval, free_dispersion, flags, graph,
                     peer_group=dist.group.WORLD if peer_ok else None, row_offset=lo)
    out = ShardedEM(be, N, kval, rank, world, dist.group.WORLD).run(beta32, itermax, 0.01)
    mu, eps, pk, _ = be.plan.get_params()
    run = engine.NemRun(out.T, mu, eps, pk, list(out.crit), out.n_iter, out.converged, out.status,
                        dev_in.n_levels if graph is not None else 1, 0.0)
    be.plan.close()
    return run


def _run_instance(dev_in, kval: int, beta: float, free_dispersion: bool, itermax: int, flags: int, want_params: bool = True,
                  names=None):
    """One NEM instance resident in HBM: row-sharded over the process group when it is large, else on this rank's GPU."""
    run = _run_sharded(dev_in, dev_in.D, names, kval, beta, free_dispersion, itermax, flags)
    if run is None:
        run = engine.nem_run_device(dev_in, kval, beta, free_dispersion, itermax=itermax, cv_th=0.01, flags=flags,
                                    want_params=want_params)
    return run


def _f32(x: float) -> float:
    """CriterT.M is a C float in the reference (nem_typ.h); the log-space epilogue sums it in fp64."""
    return float(np.float32(x))


def _g(x: float, spec: str) -> float:
    """Value after the C printf/parse round trip the reference performs through the .mf file."""
    return float(spec % x)


def run_partitioning(
    nem_dir_path: Path,
    nb_org: int,
    beta: float = 2.5,
    free_dispersion: bool = False,
    kval: int = 3,
    seed: int = 42,
    init: str = "param_file",
    keep_files: bool = False,
    itermax: int = 100,
    just_log_likelihood: bool = False,
) -> Union[Tuple[dict, None, None], Tuple[int, float, float], Tuple[dict, dict, float]]:
    """Reference: partition.py:35-274.  Same arguments and return values.

    ``nem_dir_path`` is the handle under which ``write_nem_input_files`` registered the exported instance.
    ``seed`` has no effect on this path in the reference either; ``keep_files`` is unused there.
    """
    logger = logging.getLogger("PPanGGOLiN")
    if init not in ("param_file", "init_from_old"):
        raise NotImplementedError("init='random' (RandNemAlgo, nem_alg.c:1605-1779) is not part of the B200 path; "
                                  "PPanGGOLiN's CLI never uses it")
    inp = export.lookup(nem_dir_path)
    if inp is None:
        raise FileNotFoundError(f"no exported NEM instance registered for {nem_dir_path}: call write_nem_input_files first")
    if inp.device is None:
        inp.device = _device_input(inp)
    run = _run_instance(inp.device, kval, beta, free_dispersion, itermax, _epilogue_flags(nb_org), names=inp.fam_names)
    if not run.ok:
        # the reference writes no .uf in this case and partition.py:233-265 returns empty objects
        if just_log_likelihood:
            logger.warning("A NEM run failed while estimating the optimal number of partitions "
                           "(testing a candidate K), and this candidate was skipped. "
                           "Final partitioning can still succeed.")
        else:
            logger.warning("A NEM run failed for this dataset/chunk and was skipped. In chunked mode, other chunks "
                           "can still complete and final partitioning may still succeed.")
        logger.debug(f"NEM status={run.status} kval={kval} nb_org={nb_org} beta={beta} itermax={itermax}")
        return {}, None, None

    log_likelihood = _g(_f32(run.crit[3]), "%g")     # partition.py:187 reads criterion M (a C float) from "%g" text
    labels, entropy = engine.labels_device(run.T)    # partition.py:206-231 on " %5.3f " text
    if just_log_likelihood:
        return kval, log_likelihood, entropy

    all_parameters = {}
    for k in range(kval):                            # partition.py:192-204 (" %10.3g ", "  %5.3g  ", " %10g ")
        mu_k = [bool(_g(float(v), "%10.3g")) for v in run.mu[k]]
        epsilon_k = [_g(float(v), "%10g") for v in run.eps[k]]
        proportion = _g(float(run.pk[k]), "%5.3g")
        if k == 0:
            all_parameters["persistent"] = (mu_k, epsilon_k, proportion)
        elif k == kval - 1:
            all_parameters["cloud"] = (mu_k, epsilon_k, proportion)
        else:
            all_parameters["shell_" + str(k)] = (mu_k, epsilon_k, proportion)

    parti = {0: "P", kval - 1: "C", -1: "S_"}
    for i in range(1, kval - 1):
        parti[i] = "S" + str(i)
    lab = labels.cpu().numpy()
    partitions_list = [parti[int(v)] for v in lab]
    return dict(zip(inp.fam_names, partitions_list)), all_parameters, log_likelihood


def nem_single(args) -> Union[Tuple[dict, None, None], Tuple[int, float, float], Tuple[dict, dict, float]]:
    """partition.py:277-287"""
    return run_partitioning(*args)


def partition_nem(
    index: int,
    kval: int,
    beta: float = 2.5,
    sm_degree: int = 10,
    free_dispersion: bool = False,
    seed: int = 42,
    init: str = "param_file",
    tmpdir: Path = None,
    keep_tmp_files: bool = False,
) -> Union[Tuple[dict, None, None], Tuple[int, float, float], Tuple[dict, dict, float]]:
    """partition.py:290-330"""
    currtmpdir = tmpdir / f"{str(index)}"
    samp = samples[index]
    edges_weight, nb_fam = write_nem_input_files(tmpdir=currtmpdir, organisms=samp, sm_degree=sm_degree,
                                                 keep_tmp_files=keep_tmp_files)
    try:
        return run_partitioning(currtmpdir, len(samp), beta * (nb_fam / edges_weight), free_dispersion, kval=kval,
                                seed=seed, init=init, keep_files=keep_tmp_files)
    finally:
        export.release(currtmpdir)


def nem_samples(pack: tuple):
    """partition.py:333-341"""
    return partition_nem(*pack)


_ARRAYS = {"pan": None, "arrays": None}


def _arrays_for(pangenome):
    """Device-resident array form of ``pangenome`` (export_device.DevicePangenome), built once per partition() call; None
    when it was not prepared (keep_tmp_files: the object traversal that also produces the text files is used)."""
    if _ARRAYS["pan"] is pangenome:
        return _ARRAYS["arrays"]
    return None


def prepare_arrays(pangenome, device=None):
    """Read the pangenome objects once into arrays in HBM; every later write_nem_input_files(...) for this pangenome derives
    its instance from them on the device.  Called by partition(); callers that drive write_nem_input_files themselves
    (rarefaction) call it too."""
    from .export_device import DevicePangenome
    _ARRAYS["pan"], _ARRAYS["arrays"] = pangenome, DevicePangenome.from_pangenome(pangenome, device or "cuda")
    return _ARRAYS["arrays"]


def write_nem_input_files(tmpdir: Path, organisms: set, sm_degree: int = 10, keep_tmp_files: bool = False) -> Tuple[float, int]:
    """partition.py:344-436.  Exports the sub-pangenome of ``organisms`` (global ``pan``) and registers it under
    ``tmpdir``; returns (total edge weight / 2, number of families) exactly like the reference."""
    dp = _arrays_for(pan)
    if dp is not None and not keep_tmp_files:
        cols = dp.columns(organisms)
        inst = dp.export([cols], sm_degree)[0]               # CUDA exporter on the cached array form of the pangenome
        rows = inst.fam_rows.cpu().numpy()
        inp = export.NemInput(None, inst.D, None, None, None, [dp.fam_names[i] for i in rows], [dp.org_names[c] for c in cols],
                              inst.edges_weight)
        inp.device, inp.fam_rows, inp.bits_tensor = inst.inp, inst.fam_rows, inst.inp.bits
    else:
        inp = export.export_from_pangenome(pan, organisms, sm_degree, keep_text=keep_tmp_files)
    export.register(tmpdir, inp)
    if keep_tmp_files:
        export.write_text_files(Path(tmpdir), inp)
    return inp.edges_weight, inp.N


def evaluate_nb_partitions(
    organisms: set,
    output: Path = None,
    sm_degree: int = 10,
    free_dispersion: bool = False,
    chunk_size: int = 500,
    krange: list = None,
    icl_margin: float = 0.05,
    draw_icl: bool = False,
    cpu: int = 1,
    seed: int = 42,
    tmpdir: Path = None,
    disable_bar: bool = False,
) -> int:
    """partition.py:439-634: BIC/ICL sweep over K = krange[0]-1 .. krange[1] with beta = 0 and 10 iterations."""
    tmpdir = Path(tempfile.gettempdir()) if tmpdir is None else tmpdir
    newtmpdir = tmpdir / "eval_partitions"
    dp = _arrays_for(pan)
    if dp is not None:      # the whole sweep on arrays (nem/workflow.py): same draws, same rule
        from . import workflow
        chosen_k, all_bics, all_icls, all_lls = workflow.evaluate_nb_partitions_arrays(
            dp, [dp.org_index[o] for o in organisms], free_dispersion, chunk_size, krange, icl_margin, sm_degree)
        if len(all_bics) > 0 and draw_icl:
            _draw_icl(output, all_bics, all_icls, all_lls, chosen_k)
        return chosen_k
    if len(organisms) > chunk_size:
        select_organisms = _same_on_all_ranks(set(random.sample(list(organisms), chunk_size)), organisms)
    else:
        select_organisms = set(organisms)

    _, nb_fam = write_nem_input_files(newtmpdir, select_organisms, sm_degree)
    # the K candidates are independent NEM instances over the same matrix (beta 0, 10 iterations, partition.py:482-496):
    # one batched call instead of the reference's fork pool; under torch.distributed the candidates are split over ranks
    ks = list(range(krange[0] - 1, krange[1] + 1))
    if ks and ks[-1] > engine.MAX_K:
        logging.getLogger("PPanGGOLiN").warning(f"K candidates above {engine.MAX_K} are not supported by the B200 path "
                                                f"(NEMB200_MAX_K) and are skipped.")
        ks = [k for k in ks if k <= engine.MAX_K]
    all_log_likelihood = run_k_candidates(newtmpdir, len(select_organisms), ks, free_dispersion)
    chosen_k, all_bics, all_icls, all_lls = select_k(all_log_likelihood, len(select_organisms), nb_fam,
                                                     free_dispersion, icl_margin)
    if len(all_bics) > 0 and draw_icl:
        _draw_icl(output, all_bics, all_icls, all_lls, chosen_k)
    export.release(newtmpdir)
    return chosen_k


def run_k_candidates(nem_dir_path: Path, nb_org: int, ks: List[int], free_dispersion: bool):
    """[(K, M, entropy) | ({}, None, None)] for every K in ``ks`` -- what ``nem_single`` returns with
    ``just_log_likelihood=True`` (partition.py:271-272), computed as one batch on the GPU."""
    logger = logging.getLogger("PPanGGOLiN")
    inp = export.lookup(nem_dir_path)
    if inp is None:
        raise FileNotFoundError(f"no exported NEM instance registered for {nem_dir_path}")
    if inp.device is None:
        inp.device = _device_input(inp)
    rank, world, dist = _rank_world()
    mine = ks[rank::world]
    runs = engine.nem_run_batch([(inp.device, k, 0.0, free_dispersion, 10) for k in mine], flags=_epilogue_flags(nb_org))
    results = []
    for k, run in zip(mine, runs):
        if not run.ok:
            logger.warning("A NEM run failed while estimating the optimal number of partitions "
                           "(testing a candidate K), and this candidate was skipped. Final partitioning can still succeed.")
            results.append(({}, None, None))
            continue
        _, entropy = engine.labels_device(run.T)
        results.append((k, _g(_f32(run.crit[3]), "%g"), entropy))
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = [r for part in gathered for r in part]
    return results


def select_k(all_log_likelihood, nb_org: int, nb_fam: int, free_dispersion: bool, icl_margin: float):
    """partition.py:516-545"""
    all_bics, all_icls, all_lls = defaultdict(float), defaultdict(float), defaultdict(float)
    for k_candidate, log_likelihood, entropy in all_log_likelihood:
        if log_likelihood is not None:
            nb_params = k_candidate * (nb_org + 1 + (nb_org if free_dispersion else 1))
            all_bics[k_candidate] = log_likelihood - 0.5 * (math.log(nb_params) * nb_fam)
            all_icls[k_candidate] = all_bics[k_candidate] - entropy
            all_lls[k_candidate] = log_likelihood
    chosen_k = 3
    if len(all_bics) > 3:
        max_icl_k = max(all_icls, key=all_icls.get)
        delta_icl = (all_icls[max_icl_k] - min(all_icls.values())) * icl_margin
        best_k = min({k for k, icl in all_icls.items() if icl >= all_icls[max_icl_k] - delta_icl and k <= max_icl_k})
        chosen_k = best_k if best_k >= 3 else chosen_k
    return chosen_k, all_bics, all_icls, all_lls


def _draw_icl(output, all_bics, all_icls, all_lls, best_k):
    """partition.py:547-633 (plotly figure); only available when plotly is installed."""
    try:
        import plotly.graph_objs as go
        import plotly.offline as out_plotly
    except ImportError:
        logging.getLogger("PPanGGOLiN").warning("plotly is not installed: the ICL curve is not drawn")
        return
    ks = sorted(all_bics.keys())
    traces = [go.Scatter(x=ks, y=[all_bics[k] for k in ks], name="BIC", mode="lines+markers"),
              go.Scatter(x=ks, y=[all_icls[k] for k in ks], name="ICL", mode="lines+markers"),
              go.Scatter(x=ks, y=[all_lls[k] for k in ks], name="log likelihood", mode="lines+markers")]
    layout = go.Layout(title="ICL curve (best K is " + str(best_k) + ")", titlefont=dict(size=20),
                       xaxis=dict(title="number of partitions"), yaxis=dict(title="ICL, BIC, log likelihood"),
                       plot_bgcolor="#ffffff")
    fig = go.Figure(data=traces, layout=layout)
    out_plotly.plot(fig, filename=(Path(output) / f"ICL_curve_K{str(best_k)}.html").as_posix(), auto_open=False)


def check_pangenome_former_partition(pangenome: Pangenome, force: bool = False):
    """partition.py:637-650"""
    if pangenome.status["partitioned"] == "inFile" and not force:
        raise Exception("You are trying to partition a pangenome already partitioned."
                        " If you REALLY want to do that, "
                        "use --force (it will erase partitions and every feature computed from them.")
    elif pangenome.status["partitioned"] == "inFile" and force:
        erase_pangenome(pangenome, partition=True)


def partition(
    pangenome: Pangenome,
    output: Path = None,
    beta: float = 2.5,
    sm_degree: int = 10,
    free_dispersion: bool = False,
    chunk_size: int = 500,
    kval: int = -1,
    krange: list = None,
    icl_margin: float = 0.05,
    draw_icl: bool = False,
    cpu: int = 1,
    seed: int = 42,
    tmpdir: Path = None,
    keep_tmp_files: bool = False,
    force: bool = False,
    disable_bar: bool = False,
):
    """partition.py:653-899: same parameters, same side effects on ``pangenome``."""
    tmpdir = Path(tempfile.gettempdir()) if tmpdir is None else Path(tmpdir)
    kmm = [3, 20] if krange is None else krange
    global samples
    global pan

    pan = pangenome
    logger = logging.getLogger("PPanGGOLiN")
    if draw_icl and output is None:
        raise Exception("Combination of option impossible: "
                        "You asked to draw the ICL curves but did not provide an output directory!")
    check_pangenome_former_partition(pangenome, force)
    check_pangenome_info(pangenome, need_annotations=True, need_families=True, need_graph=True, disable_bar=disable_bar)
    organisms = set(pangenome.organisms)
    dp = None
    if not keep_tmp_files:
        dp = prepare_arrays(pangenome)   # one traversal of the objects; every sample / chunk below is derived on the device

    if keep_tmp_files:
        tmp_dir = tempfile.mkdtemp(dir=tmpdir)
        tmp_path = Path(tmp_dir)
    else:
        tmp_dir = tempfile.TemporaryDirectory(dir=tmpdir)
        tmp_path = Path(tmp_dir.name)

    if len(organisms) <= 10:
        logger.warning(f"The number of selected genomes is too low ({len(organisms)} "
                       f"genomes used) to robustly partition the graph")

    pangenome.parameters["partition"] = {}
    pangenome.parameters["partition"]["beta"] = beta
    pangenome.parameters["partition"]["max_degree_smoothing"] = sm_degree
    pangenome.parameters["partition"]["free_dispersion"] = free_dispersion
    pangenome.parameters["partition"]["ICL_margin"] = icl_margin
    pangenome.parameters["partition"]["seed"] = seed
    if len(organisms) > chunk_size:
        pangenome.parameters["partition"]["chunk_size"] = chunk_size
    pangenome.parameters["partition"]["# computed nb of partitions"] = False
    pangenome.parameters["partition"]["nb_of_partitions"] = kval
    if kval < 2:
        pangenome.parameters["partition"]["# computed nb of partitions"] = True
        logger.info("Estimating the optimal number of partitions...")
        kval = evaluate_nb_partitions(organisms, output, sm_degree, free_dispersion, chunk_size, kmm, icl_margin,
                                      draw_icl, cpu, seed, tmp_path, disable_bar)
        logger.info(f"The number of partitions has been evaluated at {kval}")

    pangenome.parameters["partition"]["# final nb of partitions"] = kval
    pangenome.parameters["partition"]["krange"] = kmm
    init = "param_file"

    partitioning_results = {}
    start_partitioning = time.time()
    logger.info("Partitioning...")
    if dp is not None:
        # the workflow on arrays (nem/workflow.py): chunk samples / the single instance exported and run on the GPU in batches
        from . import workflow
        order = list(organisms)                                  # what the reference shuffles: list(set of organisms)
        outcome = workflow.partition_arrays(dp, beta, sm_degree, free_dispersion, chunk_size, kval, kmm, icl_margin, seed,
                                            cols_in_order=[dp.org_index[o] for o in order])
        by_col = {dp.org_index[o]: o for o in order}
        samples.extend({by_col[c] for c in smp} for smp in outcome.samples)
        partitioning_results = [{name: lab for name, lab in zip(dp.fam_names, outcome.labels) if lab != ""}, []]
    elif chunk_size < len(organisms):
        # keep_tmp_files: every chunk through write_nem_input_files / run_partitioning, which also write the reference's files
        families = list(pangenome.gene_families)
        pansize = len(families)
        cpt_partition = {fam.name: {"P": 0, "S": 0, "C": 0, "U": 0} for fam in families}
        random.seed(seed)
        validated = set()

        def validate_family(res):  # partition.py:784-802
            for node, nem_class in res[0].items():
                cpt_partition[node][nem_class[0]] += 1
                sum_partionning = sum(cpt_partition[node].values())
                if (sum_partionning > len(organisms) / chunk_size
                        and max(cpt_partition[node].values()) >= sum_partionning * 0.5) or (sum_partionning > len(organisms)):
                    if node not in validated:
                        if max(cpt_partition[node].values()) < sum_partionning * 0.5:
                            cpt_partition[node]["U"] = len(organisms)
                        validated.add(node)

        org_nb_sample = Counter()
        for org in organisms:
            org_nb_sample[org] = 0
        condition = len(organisms) / chunk_size
        while len(validated) < pansize:
            prev = len(samples)
            _draw_samples(organisms, chunk_size, org_nb_sample, condition)
            logger.info("Launching NEM")
            for i, _ in enumerate(samples[prev:], start=prev):
                validate_family(nem_samples((i, kval, beta, sm_degree, free_dispersion, seed, init, tmp_path, keep_tmp_files)))
            condition += 1
            logger.debug(f"There are {len(validated)} validated families out of {pansize} families.")
        for fam, data in cpt_partition.items():
            partitioning_results[fam] = max(data, key=data.get)
        partitioning_results = [partitioning_results, []]
        logger.info(f"Did {len(samples)} partitioning with chunks of size {chunk_size} among "
                    f"{len(organisms)} genomes in {round(time.time() - start_partitioning, 2)} seconds.")
    else:
        random.seed(seed)
        edges_weight, nb_fam = write_nem_input_files(tmp_path / "0", organisms, sm_degree=sm_degree, keep_tmp_files=keep_tmp_files)
        partitioning_results = run_partitioning(tmp_path / "0", len(organisms), beta * (nb_fam / edges_weight), free_dispersion,
                                                kval=kval, seed=seed, init=init, keep_files=keep_tmp_files)
        if partitioning_results == [{}, None, None]:  # never true in the reference either (tuple vs list)
            raise Exception("Statistical partitioning does not work on your data. "
                            "This usually happens because you used very few (<15) genomes.")
        logger.info(f"Partitioned {len(organisms)} genomes in {round(time.time() - start_partitioning, 2)} seconds.")

    for fam_name, part in partitioning_results[0].items():
        pangenome.get_gene_family(fam_name).partition = part

    pangenome.status["partitioned"] = "Computed"
    export.release(tmp_path)
    _ARRAYS["pan"], _ARRAYS["arrays"] = None, None
    if not keep_tmp_files:
        tmp_dir.cleanup()
    else:
        copytree(tmp_path, output / "NEM_files/")


def _draw_samples(organisms, chunk_size, org_nb_sample, condition):
    """partition.py:810-818: shuffle the genomes and cut chunks until every genome was drawn `condition` times.
    The strict '>' drops the tail of every shuffle."""
    while not all(val >= condition for val in org_nb_sample.values()):
        shuffled_orgs = list(organisms)
        random.shuffle(shuffled_orgs)
        while len(shuffled_orgs) > chunk_size:
            samples.append(set(shuffled_orgs[:chunk_size]))
            for org in samples[-1]:
                org_nb_sample[org] += 1
            shuffled_orgs = shuffled_orgs[chunk_size:]


def _same_on_all_ranks(chosen: set, organisms) -> set:
    """Random genome samples must be the same on every rank of a process group, but they are drawn from sets of objects whose
    iteration order depends on object addresses: rank 0's draw is broadcast by genome name."""
    rank, world, dist = _rank_world()
    if world == 1:
        return chosen
    payload = [sorted(o.name for o in chosen) if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    if rank == 0:
        return chosen
    by_name = {o.name: o for o in organisms}
    return {by_name[n] for n in payload[0]}


_local_only = 0


class local_instances:
    """Context in which this module does not deal work out to the process group: used by callers that already run
    different jobs on different ranks (rarefaction samples), where a collective inside would not be matched."""

    def __enter__(self):
        global _local_only
        _local_only += 1
        return self

    def __exit__(self, *exc):
        global _local_only
        _local_only -= 1
        return False


def _rank_world():
    """(rank, world, torch.distributed or None): chunk samples and K candidates are independent NEM instances and are dealt
    out to the ranks of an initialised process group ("independent instances")."""
    if _local_only:
        return 0, 1, None
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), dist
    except ImportError:
        pass
    return 0, 1, None


def launch(args: argparse.Namespace):
    """partition.py:902-932"""
    if args.draw_ICL or args.keep_tmp_files:
        mk_outdir(args.output, args.force)
    global pan
    pan.add_file(args.pangenome)
    partition(pan, args.output, args.beta, args.max_degree_smoothing, args.free_dispersion, args.chunk_size,
              args.nb_of_partitions, args.krange, args.ICL_margin, args.draw_ICL, args.cpu, args.seed, args.tmpdir,
              args.keep_tmp_files, args.force, disable_bar=args.disable_prog_bar)
    logging.getLogger("PPanGGOLiN").debug("Write partition in pangenome")
    write_pangenome(pan, pan.file, args.force, disable_bar=args.disable_prog_bar)
    logging.getLogger("PPanGGOLiN").debug("Partitioning is finished")


def subparser(sub_parser: argparse._SubParsersAction) -> argparse.ArgumentParser:
    """partition.py:935-947"""
    parser = sub_parser.add_parser("partition", formatter_class=argparse.RawTextHelpFormatter)
    parser_partition(parser)
    return parser


def parser_partition(parser: argparse.ArgumentParser):
    """partition.py:950-1078: same flags, defaults and types."""
    required = parser.add_argument_group(title="Required arguments", description="One of the following arguments is required :")
    required.add_argument("-p", "--pangenome", required=False, type=Path, help="The pangenome.h5 file")
    optional = parser.add_argument_group(title="Optional arguments")
    optional.add_argument("-b", "--beta", required=False, default=2.5, type=float,
                          help="beta is the strength of the smoothing using the graph topology during partitioning. "
                               "0 will deactivate spatial smoothing.")
    optional.add_argument("-ms", "--max_degree_smoothing", required=False, default=10, type=float,
                          help="max. degree of the nodes to be included in the smoothing process.")
    optional.add_argument("-o", "--output", required=False, type=Path,
                          default=Path(f"ppanggolin_output{time.strftime('DATE%Y-%m-%d_HOUR%H.%M.%S', time.localtime())}"
                                       f"_PID{str(os.getpid())}"), help="Output directory")
    optional.add_argument("-fd", "--free_dispersion", required=False, default=False, action="store_true",
                          help="use if the dispersion around the centroid vector of each partition during must be free.")
    optional.add_argument("-ck", "--chunk_size", required=False, default=500, type=int,
                          help="Size of the chunks when performing partitioning using chunks of genomes.")
    optional.add_argument("-K", "--nb_of_partitions", required=False, default=-1, type=int,
                          help="Number of partitions to use. Must be at least 2. If under 2, it will be detected automatically.")
    optional.add_argument("-Kmm", "--krange", nargs=2, required=False, type=int, default=[3, 20],
                          help="Range of K values to test when detecting K automatically.")
    optional.add_argument("-im", "--ICL_margin", required=False, type=float, default=0.05,
                          help="K is detected automatically by maximizing ICL. The smallest K within this margin of the "
                               "maximal ICL is kept.")
    optional.add_argument("--draw_ICL", required=False, default=False, action="store_true",
                          help="Use if you want to draw the ICL curve for all the tested K values.")
    optional.add_argument("--keep_tmp_files", required=False, default=False, action="store_true",
                          help="Use if you want to keep the temporary NEM files")
    optional.add_argument("-se", "--seed", type=int, default=42, help="seed used to generate random numbers")
    optional.add_argument("-c", "--cpu", required=False, default=1, type=int, help="Number of available cpus")
    optional.add_argument("--tmpdir", required=False, type=Path, default=Path(tempfile.gettempdir()),
                          help="directory for storing temporary files")
