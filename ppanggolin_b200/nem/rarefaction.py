#!/usr/bin/env python3
"""Second caller of the NEM API: rarefaction (SURVEY.md section 8(f)-2;
/root/reference/ppanggolin/nem/rarefaction.py:36-203 ``raref_nem`` / ``launch_raref_nem``, :538-724
``make_rarefaction_curve``).

Same names, arguments and results for the partitioning part: genome samples of growing size, per sample the exact /
soft core and accessory counts and the NEM partition counts, written to ``rarefaction.csv`` with the reference's
columns (rarefaction.py:214-252).  Differences, all below the API:

* no fork pool (rarefaction.py:707): a CUDA context does not survive ``fork``; with a fixed K the samples of one size are
  exported and run as batches (one CUDA exporter call + one batched EM call per 16 samples, ``raref_batch``), the others
  one after the other, every instance derived on the device from the array form of the pangenome (``export_device``);
* the per-sample core / soft-core statistics come from a popcount kernel over the bit-packed presence matrix
  (``nemb200_subset_counts``) instead of gmpy2 bitarrays (rarefaction.py:657-684) -- same integer counts;
* ``draw_curve`` (rarefaction.py:205-535) writes ``rarefaction.csv`` and the Heaps'-law fit ``rarefaction_parameters.csv``
  (scipy); the plotly figure is produced only when plotly is importable -- neither is part of the hot path;
* ``launch`` / ``subparser`` / ``parser_rarefaction`` (rarefaction.py:727-912): same flags, defaults and types.
"""
from __future__ import annotations

import argparse
import logging
import os
import random
import tempfile
import time
from collections import Counter
from pathlib import Path
from typing import Dict, List, Tuple, Union

from . import partition as ppp

samples: List[set] = []

RAREF_BATCH = 16          # samples per exporter / EM call

CSV_COLUMNS = ["genomes_count", "persistent", "shell", "cloud", "undefined", "exact_core", "exact_accessory", "soft_core",
               "soft_accessory", "pangenome", "K"]


def raref_nem(index: int, tmpdir: Path, beta: float = 2.5, sm_degree: int = 10, free_dispersion: bool = False,
              chunk_size: int = 500, kval: int = -1, krange: list = None, seed: int = 42) -> Tuple[Dict[str, int], int]:
    """rarefaction.py:36-189: partition counts of sample ``index`` (chunked voting when the sample exceeds chunk_size)."""
    samp = samples[index]
    currtmpdir = Path(tmpdir) / f"{str(index)}"
    kmm = [3, 20] if krange is None else krange
    if kval < 3:
        kval = ppp.evaluate_nb_partitions(organisms=samp, sm_degree=sm_degree, free_dispersion=free_dispersion,
                                          chunk_size=chunk_size, krange=kmm, seed=seed, tmpdir=Path(tmpdir) / f"{str(index)}_eval")
    if len(samp) <= chunk_size:
        edges_weight, nb_fam = ppp.write_nem_input_files(tmpdir=currtmpdir, organisms=set(samp), sm_degree=sm_degree)
        try:
            cpt_partition = ppp.run_partitioning(currtmpdir, len(samp), beta * (nb_fam / edges_weight), free_dispersion,
                                                 kval=kval, seed=seed, init="param_file")[0]
        finally:
            ppp.export.release(currtmpdir)
    else:
        families = set()
        cpt_partition = {}
        validated = set()
        cpt = 0

        def validate_family(result):   # rarefaction.py:100-122
            for node, nem_class in result[0].items():
                cpt_partition[node][nem_class[0]] += 1
                sum_partitioning = sum(cpt_partition[node].values())
                if (sum_partitioning > len(samp) / chunk_size and max(cpt_partition[node].values()) >= sum_partitioning * 0.5) \
                        or (sum_partitioning > len(samp)):
                    if node not in validated:
                        if max(cpt_partition[node].values()) < sum_partitioning * 0.5:
                            cpt_partition[node]["U"] = len(samp)
                        validated.add(node)

        for fam in ppp.pan.gene_families:
            if not samp.isdisjoint(set(fam.organisms)):
                families.add(fam)
                cpt_partition[fam.name] = {"P": 0, "S": 0, "C": 0, "U": 0}
        org_nb_sample = Counter()
        for org in samp:
            org_nb_sample[org] = 0
        condition = len(samp) / chunk_size
        while len(validated) < len(families):
            org_samples = []
            while not all(val >= condition for val in org_nb_sample.values()):
                shuffled_orgs = list(samp)
                random.shuffle(shuffled_orgs)
                while len(shuffled_orgs) > chunk_size:
                    org_samples.append(set(shuffled_orgs[:chunk_size]))
                    for org in org_samples[-1]:
                        org_nb_sample[org] += 1
                    shuffled_orgs = shuffled_orgs[chunk_size:]
            for sub in org_samples:
                d = currtmpdir / f"{str(cpt)}"
                edges_weight, nb_fam = ppp.write_nem_input_files(d, sub, sm_degree=sm_degree)
                try:
                    validate_family(ppp.run_partitioning(d, len(sub), beta * (nb_fam / edges_weight), free_dispersion,
                                                         kval=kval, seed=seed, init="param_file"))
                finally:
                    ppp.export.release(d)
                cpt += 1
            condition += 1   # the reference loops forever if one pass does not validate every family; draw more instead
    if len(cpt_partition) == 0:
        counts = {"persistent": "NA", "shell": "NA", "cloud": "NA", "undefined": "NA", "K": kval}
    else:
        counts = {"persistent": 0, "shell": 0, "cloud": 0, "undefined": 0, "K": kval}
        for val in cpt_partition.values():
            part = val if isinstance(val, str) else max(val, key=val.get)
            if part.startswith("P"):
                counts["persistent"] += 1
            elif part.startswith("C"):
                counts["cloud"] += 1
            elif part.startswith("S"):
                counts["shell"] += 1
            else:
                counts["undefined"] += 1
    return counts, index


def launch_raref_nem(args) -> Tuple[Dict[str, int], int]:
    """rarefaction.py:192-203"""
    return raref_nem(*args)


def sample_family_stats(dp, all_samples: List[set], soft_core: float = 0.95) -> List[Counter]:
    """rarefaction.py:657-684: per sample, families present in all / in >= soft_core of the sample's genomes (exact / soft
    core) and the others that occur at least once (exact / soft accessory).  The reference ANDs gmpy2 bitarrays per family;
    here one popcount kernel over the bit-packed presence matrix gives every family's count in the sample
    (``nemb200_subset_counts``)."""
    import torch
    out = []
    for samp in all_samples:
        nb = dp.subset_counts(dp.columns(samp)).to(torch.int64)
        present = nb > 0
        n = len(samp)
        soft = present & (nb.to(torch.float64) >= n * soft_core)
        stats = torch.stack([(nb == n).sum(), (present & (nb != n)).sum(), soft.sum(), (present & ~soft).sum()]).cpu().tolist()
        part = Counter()
        part["nborgs"] = n
        part["exact_core"], part["exact_accessory"], part["soft_core"], part["soft_accessory"] = (int(v) for v in stats)
        out.append(part)
    return out


def _counts_from_labels(lab, kval: int) -> Dict[str, int]:
    """Partition counts of one un-chunked run (rarefaction.py:163-177): label 0 = persistent, K - 1 = cloud, the rest
    (shell classes and the undecided "S_") = shell."""
    import torch
    n = int(lab.numel())
    pers = int((lab == 0).sum().item())
    cloud = int((lab == kval - 1).sum().item())
    return {"persistent": pers, "shell": n - pers - cloud, "cloud": cloud, "undefined": 0, "K": kval}


def raref_batch(dp, indices: List[int], beta: float, sm_degree: float, free_dispersion: bool, kval: int) -> Dict[int, Dict[str, int]]:
    """The samples ``indices`` (all of the same size, none above the chunk size, K fixed) as ONE exporter call and ONE batched
    EM call instead of one fork-pool task each (rarefaction.py:707-716)."""
    from . import engine
    insts = dp.export([dp.columns(samples[i]) for i in indices], sm_degree)
    jobs = [(it.inp, kval, it.beta_eff(beta), free_dispersion, 100) for it in insts]
    runs = engine.nem_run_batch(jobs, flags=ppp._epilogue_flags(insts[0].D), want_params=False)
    out = {}
    for i, run in zip(indices, runs):
        if not run.ok:
            out[i] = {"persistent": "NA", "shell": "NA", "cloud": "NA", "undefined": "NA", "K": kval}
            continue
        lab, _ = engine.labels_device(run.T)
        out[i] = _counts_from_labels(lab, kval)
    return out


def write_rarefaction_csv(output: Path, data: list) -> Path:
    """rarefaction.py:214-252"""
    name = Path(output) / "rarefaction.csv"
    with open(name, "w") as raref:
        raref.write(",".join(CSV_COLUMNS) + "\n")
        for part in data:
            raref.write(",".join(map(str, [part["nborgs"], part["persistent"], part["shell"], part["cloud"], part["undefined"],
                                           part["exact_core"], part["exact_accessory"], part["soft_core"],
                                           part["soft_accessory"], part["exact_core"] + part["exact_accessory"], part["K"]])) + "\n")
    return name


def make_rarefaction_curve(pangenome, output: Path, tmpdir: Path = None, beta: float = 2.5, depth: int = 30,
                           min_sampling: int = 1, max_sampling: int = 100, sm_degree: int = 10, free_dispersion: bool = False,
                           chunk_size: int = 500, kval: int = -1, krange: list = None, cpu: int = 1, seed: int = 42,
                           kestimate: bool = False, soft_core: float = 0.95, disable_bar: bool = False):
    """rarefaction.py:538-724, same parameters; returns the per-sample records that go into ``rarefaction.csv``."""
    logger = logging.getLogger("PPanGGOLiN")
    tmpdir = Path(tempfile.gettempdir()) if tmpdir is None else Path(tmpdir)
    if krange is None:
        krange = [3, -1]
    ppp.pan = pangenome
    try:
        final_k = ppp.pan.parameters["partition"]["# final nb of partitions"]
        krange = [final_k if krange[0] < 0 else krange[0], final_k if krange[1] < 0 else krange[1]]
    except KeyError:
        krange = [3, 20]
    ppp.check_pangenome_info(pangenome, need_annotations=True, need_families=True, need_graph=True, disable_bar=disable_bar)
    tmpdir_obj = tempfile.TemporaryDirectory(dir=tmpdir)
    tmp_path = Path(tmpdir_obj.name)
    organisms = list(pangenome.organisms)
    max_sampling = len(organisms) if float(len(organisms)) < max_sampling else int(max_sampling)
    arrays = ppp.prepare_arrays(pangenome)   # one traversal of the objects for every sample below

    if kval < 3 and kestimate is False:
        try:
            kval = ppp.pan.parameters["partition"]["# final nb of partitions"]
            logger.info(f"Reuse the number of partitions {kval}")
        except KeyError:
            logger.info("Estimating the number of partitions...")
            kval = ppp.evaluate_nb_partitions(organisms=set(organisms), sm_degree=sm_degree, free_dispersion=free_dispersion,
                                              chunk_size=chunk_size, krange=krange, cpu=cpu, seed=seed, tmpdir=tmp_path)
            logger.info(f"The number of partitions has been evaluated at {kval}")

    logger.info("Extracting samples ...")
    rank, world, dist = ppp._rank_world()
    all_samples = []
    if rank == 0:
        for i in range(min_sampling, max_sampling):
            for _ in range(depth):
                all_samples.append(set(random.sample(organisms, i + 1)))
    if world > 1:   # the same samples on every rank: rank 0's draw, by genome name (object order differs between processes)
        payload = [[sorted(o.name for o in smp) for smp in all_samples] if rank == 0 else None]
        dist.broadcast_object_list(payload, src=0)
        if rank != 0:
            by_name = {o.name: o for o in organisms}
            all_samples = [{by_name[n] for n in names} for names in payload[0]]
    logger.info(f"Done sampling genomes in the pan, there are {len(all_samples)} samples")

    samp_nb_per_part = sample_family_stats(arrays, all_samples, soft_core)

    global samples
    samples = all_samples
    logger.info(" Partitioning all samples...")
    # the samples are independent jobs: every rank takes its share (round-robin) and runs it without collectives inside,
    # the per-sample counts are gathered at the end (the reference spreads them over a fork pool, rarefaction.py:707)
    mine = {}
    with ppp.local_instances():
        my_indices = list(range(rank, len(samples), world))
        if kval >= 3:
            # K is fixed: samples of one size that fit a single NEM instance are exported and run as batches
            by_size = {}
            for index in my_indices:
                if len(samples[index]) <= chunk_size:
                    by_size.setdefault(len(samples[index]), []).append(index)
            for size, idxs in by_size.items():
                for b0 in range(0, len(idxs), RAREF_BATCH):
                    mine.update(raref_batch(arrays, idxs[b0:b0 + RAREF_BATCH], beta, sm_degree, free_dispersion, kval))
        for index in my_indices:
            if index in mine:
                continue
            counts, idx = launch_raref_nem((index, tmp_path, beta, sm_degree, free_dispersion, chunk_size, kval, krange, seed))
            mine[idx] = counts
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        mine = {k: v for part in gathered for k, v in part.items()}
    for idx, counts in mine.items():
        samp_nb_per_part[idx] = {**counts, **samp_nb_per_part[idx]}
    logger.info("Done  partitioning everything")
    if output is not None and rank == 0:
        Path(output).mkdir(parents=True, exist_ok=True)
        write_rarefaction_csv(Path(output), samp_nb_per_part)
    tmpdir_obj.cleanup()
    ppp._ARRAYS["pan"], ppp._ARRAYS["arrays"] = None, None
    logger.info("Done making the rarefaction curves")
    return samp_nb_per_part


def draw_curve(output: Path, data: list, max_sampling: int = 10):
    """rarefaction.py:205-535: ``rarefaction.csv``, the Heaps'-law fit kappa * n^gamma of every partition on the samples of
    more than 15 genomes (``rarefaction_parameters.csv``: partition, kappa, gamma, their standard errors, IQR area) and,
    when plotly is available, ``rarefaction_curve.html``."""
    import numpy
    from scipy import optimize as optimization
    output = Path(output)
    write_rarefaction_csv(output, data)
    cols = {c: numpy.array([float("nan") if part.get(k, "NA") == "NA" else float(part[k]) for part in data])
            for c, k in (("genomes_count", "nborgs"), ("persistent", "persistent"), ("shell", "shell"), ("cloud", "cloud"),
                         ("undefined", "undefined"), ("exact_core", "exact_core"), ("exact_accessory", "exact_accessory"),
                         ("soft_core", "soft_core"), ("soft_accessory", "soft_accessory"))}
    cols["pangenome"] = cols["exact_core"] + cols["exact_accessory"]

    def heap_law(n, p_kappa, p_gamma):
        return p_kappa * n ** p_gamma

    def poly_area(p_x, p_y):
        return 0.5 * numpy.abs(numpy.dot(p_x, numpy.roll(p_y, 1)) - numpy.dot(p_y, numpy.roll(p_x, 1)))

    nb_org_min_fitting = 15
    fits = {}
    with open(output / "rarefaction_parameters.csv", "w") as params_file:
        params_file.write("partition,kappa,gamma,kappa_std_error,gamma_std_error,IQR_area\n")
        for partition in ["persistent", "shell", "cloud", "undefined", "exact_core", "exact_accessory", "soft_core", "soft_accessory",
                          "pangenome"]:
            sizes = [i for i in range(1, max_sampling + 1) if numpy.any(cols["genomes_count"] == i)]
            q25 = [numpy.nanpercentile(cols[partition][cols["genomes_count"] == i], 25) for i in sizes]
            q75 = [numpy.nanpercentile(cols[partition][cols["genomes_count"] == i], 75) for i in sizes]
            area_iqr = poly_area(sizes + list(reversed(sizes)), q25 + q75) if sizes else 0.0
            sel = (cols["genomes_count"] > nb_org_min_fitting) & ~numpy.isnan(cols[partition])
            try:
                res = optimization.curve_fit(heap_law, cols["genomes_count"][sel], cols[partition][sel], numpy.array([0.0, 0.0]))
                kappa, gamma = res[0]
                error_k, error_g = numpy.sqrt(numpy.diag(res[1]))
                if numpy.isinf(error_k) and numpy.isinf(error_g):
                    params_file.write(",".join([partition, "NA", "NA", "NA", "NA", str(area_iqr)]) + "\n")
                else:
                    params_file.write(",".join([partition, str(kappa), str(gamma), str(error_k), str(error_g), str(area_iqr)]) + "\n")
                    fits[partition] = (kappa, gamma)
            except (TypeError, RuntimeError, ValueError):  # too few points, no convergence
                params_file.write(",".join([partition, "NA", "NA", "NA", "NA", str(area_iqr)]) + "\n")
    try:
        import plotly.graph_objs as go
        import plotly.offline as out_plotly
    except ImportError:
        logging.getLogger("PPanGGOLiN").warning("plotly is not installed: rarefaction_curve.html is not drawn")
        return fits
    traces = []
    for partition, (kappa, gamma) in fits.items():
        xs = list(range(nb_org_min_fitting + 1, max_sampling + 1))
        traces.append(go.Scatter(x=xs, y=[heap_law(x, kappa, gamma) for x in xs], name=partition + ": Heaps' law", mode="lines"))
        traces.append(go.Scatter(x=cols["genomes_count"].tolist(), y=cols[partition].tolist(), name=partition, mode="markers"))
    layout = go.Layout(title="Rarefaction curve", xaxis=dict(title="size of genome subsets"), yaxis=dict(title="# of gene families"),
                       plot_bgcolor="#ffffff")
    out_plotly.plot(go.Figure(data=traces, layout=layout), filename=(output / "rarefaction_curve.html").as_posix(), auto_open=False)
    return fits


def launch(args: argparse.Namespace):
    """rarefaction.py:727-755"""
    ppp.mk_outdir(args.output, args.force)
    pangenome = ppp.Pangenome()
    pangenome.add_file(args.pangenome)
    make_rarefaction_curve(pangenome=pangenome, output=args.output, tmpdir=args.tmpdir, beta=args.beta, depth=args.depth,
                           min_sampling=args.min, max_sampling=args.max, sm_degree=args.max_degree_smoothing,
                           free_dispersion=args.free_dispersion, chunk_size=args.chunk_size, kval=args.nb_of_partitions,
                           krange=args.krange, cpu=args.cpu, seed=args.seed, kestimate=args.reestimate_K, soft_core=args.soft_core,
                           disable_bar=args.disable_prog_bar)


def subparser(sub_parser: argparse._SubParsersAction) -> argparse.ArgumentParser:
    """rarefaction.py:758-772"""
    parser = sub_parser.add_parser("rarefaction", description="Compute the rarefaction curve of the pangenome",
                                   formatter_class=argparse.RawTextHelpFormatter)
    parser_rarefaction(parser)
    return parser


def parser_rarefaction(parser: argparse.ArgumentParser):
    """rarefaction.py:775-912: same flags, defaults and types."""
    required = parser.add_argument_group(title="Required arguments", description="One of the following arguments is required :")
    required.add_argument("-p", "--pangenome", required=False, type=Path, help="The pangenome .h5 file")
    optional = parser.add_argument_group(title="Optional arguments")
    optional.add_argument("-b", "--beta", required=False, default=2.5, type=float,
                          help="beta is the strength of the smoothing using the graph topology during  partitioning. "
                               "0 will deactivate spatial smoothing.")
    optional.add_argument("--depth", required=False, default=30, type=int, help="Number of samplings at each sampling point")
    optional.add_argument("--min", required=False, default=1, type=int, help="Minimum number of genomes in a sample")
    optional.add_argument("--max", required=False, type=float, default=100,
                          help="Maximum number of genomes in a sample (if above the number of provided genomes, "
                               "the provided genomes will be the maximum)")
    optional.add_argument("-ms", "--max_degree_smoothing", required=False, default=10, type=float,
                          help="max. degree of the nodes to be included in the smoothing process.")
    optional.add_argument("-o", "--output", required=False, type=Path,
                          default=Path(f"ppanggolin_output{time.strftime('DATE%Y-%m-%d_HOUR%H.%M.%S', time.localtime())}"
                                       f"_PID{str(os.getpid())}"), help="Output directory")
    optional.add_argument("-fd", "--free_dispersion", required=False, default=False, action="store_true",
                          help="use if the dispersion around the centroid vector of each partition during must be free."
                               " It will be the same for all genomes by default.")
    optional.add_argument("-ck", "--chunk_size", required=False, default=500, type=int,
                          help="Size of the chunks when performing partitioning using chunks of genomes. "
                               "Chunk partitioning will be used automatically if the number of genomes is above this number.")
    optional.add_argument("-K", "--nb_of_partitions", required=False, default=-1, type=int,
                          help="Number of partitions to use. Must be at least 2. By default reuse K if it exists else compute it.")
    optional.add_argument("--reestimate_K", required=False, action="store_true",
                          help=" Will recompute the number of partitions for each sample "
                               "(between the values provided by --krange) (VERY intensive. Can take a long time.)")
    optional.add_argument("-Kmm", "--krange", nargs=2, required=False, type=int, default=[3, -1],
                          help="Range of K values to test when detecting K automatically. "
                               "Default between 3 and the K previously computed if there is one, or 20 if there are none.")
    optional.add_argument("--soft_core", required=False, type=float, default=0.95, help="Soft core threshold")
    optional.add_argument("-se", "--seed", type=int, default=42, help="seed used to generate random numbers")
    optional.add_argument("-c", "--cpu", required=False, default=1, type=int, help="Number of available cpus")
    optional.add_argument("--tmpdir", required=False, type=str, default=Path(tempfile.gettempdir()),
                          help="directory for storing temporary files")
