"""Command-line front end with the reference's contract (scUnif.py:60-249):
the same short and long flags, the same validation messages and exit codes, the
same log lines and the same `<outname>*.csv` output layout (np.savetxt, ',').

    python scUnif.py -sc cells.csv -ctype types.csv -bk bulk.csv -K 3 -outdir out/
    torchrun --nproc-per-node 8 scUnif.py ...        # cells / samples sharded over GPUs

Additions (not in the reference): `-seed` (the reference is unseeded).  Under
torchrun every rank parses the same files, keeps its block of rows, and rank 0
writes the outputs.
"""
import argparse
import datetime
import logging
import os
import sys
import time

import numpy as np

from . import csvio, dist
from .gem import LogitNormalGEM


def mtx2csv(filename, nparray):
    """scUnif.py:85-87: np.savetxt(handle, nparray, delimiter=',') -- same bytes, written
    by the multi-threaded native formatter (the L x N exp_S matrix is the bulk of a run's
    output)."""
    csvio.savetxt_csv(filename, nparray)


def gem2csv(dirname, gem, prefix=""):
    """scUnif.py:69-82: est_A, exp_S, est_pkappa, est_kappa, est_ptau, est_tau,
    est_alpha, exp_W.  Per-cell / per-sample outputs are gathered over ranks."""
    prefix = dirname + "/" + prefix
    st = gem.gathered_stats()
    if dist.rank() != 0:
        return
    mtx2csv(prefix + "est_A.csv", gem.A)
    if gem.hasSC:
        mtx2csv(prefix + "exp_S.csv", st["exp_S"])
        mtx2csv(prefix + "est_pkappa.csv", gem.pkappa)
        mtx2csv(prefix + "est_kappa.csv", st["exp_kappa"].transpose())
        mtx2csv(prefix + "est_ptau.csv", gem.ptau)
        mtx2csv(prefix + "est_tau.csv", st["exp_tau"].transpose())
    if gem.hasBK:
        mtx2csv(prefix + "est_alpha.csv", gem.alpha)
        mtx2csv(prefix + "exp_W.csv", st["exp_W"])


def load_from_file(filename, dtype=float, delimiter=","):
    """scUnif.py:90-94."""
    if filename is None:
        return None
    if dtype is float and delimiter == ",":
        return csvio.loadtxt_csv(filename)   # same values as np.loadtxt, parsed in parallel
    return np.loadtxt(filename, dtype=dtype, delimiter=delimiter)


def build_parser():
    """scUnif.py:103-140 -- identical flag names, types and defaults."""
    p = argparse.ArgumentParser()
    p.add_argument("-sc", "--single_cell_expr_file", type=str, default=None)
    p.add_argument("-bk", "--bulk_expr_file", type=str, default=None)
    p.add_argument("-ctype", "--single_cell_type_file", type=str, default=None)
    p.add_argument("-K", "--number_of_cell_types", type=int, default=3)
    p.add_argument("-iMarkers", "--iMarkers_file", type=str, default=None)
    p.add_argument("-init_A", "--initial_A_file", type=str, default=None)
    p.add_argument("-min_A", "--mininimal_A", type=float, default=1e-6)
    p.add_argument("-init_alpha", "--initial_alpha_file", type=str, default=None)
    p.add_argument("-no_est_alpha", "--no_est_alpha", dest="est_alpha", action="store_false")
    p.set_defaults(est_alpha=True)
    p.add_argument("-pkappa", "--initial_kappa_mean_var", nargs=2, type=float, action="store",
                   default=None)
    p.add_argument("-ptau", "--initial_tau_mean_var", nargs=2, type=float, action="store",
                   default=None)
    p.add_argument("-burnin", "--burn_in_length", type=int, default=50)
    p.add_argument("-sample", "--gibbs_sample_number", type=int, default=50)
    p.add_argument("-thin", "--gibbs_thinning", type=int, default=1)
    p.add_argument("-no_mean_approx", "--no_mean_approx", dest="bk_mean_approx",
                   action="store_false")
    p.add_argument("-mean_approx", "--mean_approx", dest="bk_mean_approx", action="store_true")
    p.set_defaults(bk_mean_approx=True)
    p.add_argument("-MLE_CONV", "--Mstep_convergence_tol", type=float, default=1e-6)
    p.add_argument("-EM_CONV", "--EM_convergence_tol", type=float, default=1e-6)
    p.add_argument("-MLE_maxiter", "--Mstep_maxiter", type=int, default=500)
    p.add_argument("-EM_maxiter", "--EM_maxiter", type=int, default=50)
    p.add_argument("-log", "--logging_file", type=str, default="gem_log.log")
    p.add_argument("-outdir", "--output_directory", type=str, default="out/")
    p.add_argument("-outname", "--output_prefix", type=str, default="gemout_")
    p.add_argument("-verbose", "--verbose_level", type=int, default=1)
    p.add_argument("-seed", "--random_seed", type=int, default=None)  # extension
    return p


def _init_distributed():
    """One process per GPU when launched by torchrun; a plain run stays single."""
    if int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return False
    import torch
    import torch.distributed as td
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    td.init_process_group(backend="nccl")
    return True


def main(argv=None):
    args = build_parser().parse_args(argv)
    distributed = _init_distributed()
    is_root = dist.rank() == 0

    if args.verbose_level <= 0:
        level = logging.ERROR
    elif args.verbose_level == 1:
        level = logging.INFO
    else:
        level = logging.DEBUG
    if not is_root:
        level = logging.ERROR
    logdir = os.path.dirname(args.logging_file)
    if is_root and len(logdir) > 0 and not os.path.exists(logdir):
        os.makedirs(logdir)
    root = logging.getLogger("")
    for h in list(root.handlers):
        root.removeHandler(h)
    if is_root:
        logging.basicConfig(level=level, filename="%s" % args.logging_file, format="%(message)s",
                            filemode="w")
    else:
        logging.basicConfig(level=level, format="%(message)s")
    console = logging.StreamHandler()
    console.setLevel(level)
    console.setFormatter(logging.Formatter("%(message)s"))
    root.addHandler(console)

    header_info = "#" * 80 + "\n"
    header_info += "Gibbs-EM for %d cell types.\n" % args.number_of_cell_types
    header_info += "Date and time: " + str(datetime.datetime.today()) + "\n"
    header_info += "Algorithm arguments:\n"
    for (argname, argvalue) in vars(args).items():
        header_info += "\t--" + argname + ": " + str(argvalue) + "\n"
    header_info += "#" * 80
    logging.info(header_info)

    logging.info("Loading data ...")
    SCexpr = load_from_file(args.single_cell_expr_file)
    BKexpr = load_from_file(args.bulk_expr_file)
    G = load_from_file(args.single_cell_type_file, dtype=int)
    init_A = load_from_file(args.initial_A_file)
    init_alpha = load_from_file(args.initial_alpha_file)
    iMarkers = load_from_file(args.iMarkers_file, dtype=int)
    K = args.number_of_cell_types
    if init_A is not None and len(init_A.shape) == 1:
        init_A = init_A[:, np.newaxis]
    if iMarkers is not None and iMarkers.ndim == 1:
        iMarkers = iMarkers[np.newaxis, :]
    if G is not None:
        G = np.atleast_1d(G)

    # validation: messages and exit code as scUnif.py:188-219
    if SCexpr is None and BKexpr is None:
        logging.error("ERROR: Must provide at least one of single cell or bulk data.")
        sys.exit(1)
    elif SCexpr is not None and BKexpr is not None and SCexpr.shape[1] != BKexpr.shape[1]:
        logging.error("ERROR: Single cell and bulk data must have same number of genes.")
        sys.exit(1)
    if SCexpr is not None:
        if G is None:
            logging.error("ERROR: Must provide cell type information for single cells.")
            sys.exit(1)
        elif SCexpr.shape[0] != G.shape[0]:
            logging.error("ERROR: Mismatched cell dimensions in `%s` and `%s`",
                          args.single_cell_expr_file, args.single_cell_type_file)
            sys.exit(1)
        elif len(set(G.tolist()) - set(range(K))) > 0:
            logging.error("ERROR: Cell types in `%s` can only take values in {0, ..., K-1}.",
                          args.single_cell_type_file)
            sys.exit(1)
    if BKexpr is not None:
        logging.info("%d bulk samples on %d genes are loaded.", BKexpr.shape[0], BKexpr.shape[1])
    if SCexpr is not None:
        logging.info("%d single cells on %d genes are loaded.\n", SCexpr.shape[0],
                     SCexpr.shape[1])
    if iMarkers is not None and np.max(iMarkers[:, 1]) > K:  # sic: `> K` as in the reference
        logging.error("ERROR: cell types in `%s` can only take values in {0, ..., K-1}.",
                      args.iMarkers_file)
        sys.exit(1)

    logging.info("Gibbs-EM started ...")
    start = time.time()
    myGEM = LogitNormalGEM(
        BKexpr=BKexpr, SCexpr=SCexpr, G=G, K=K, iMarkers=iMarkers, init_A=init_A,
        min_A=args.mininimal_A, init_alpha=init_alpha, est_alpha=args.est_alpha,
        init_pkappa=args.initial_kappa_mean_var, init_ptau=args.initial_tau_mean_var,
        burnin=args.burn_in_length, sample=args.gibbs_sample_number, thin=args.gibbs_thinning,
        bk_mean_approx=args.bk_mean_approx, MLE_CONV=args.Mstep_convergence_tol,
        MLE_maxiter=args.Mstep_maxiter, EM_CONV=args.EM_convergence_tol,
        EM_maxiter=args.EM_maxiter, seed=args.random_seed)
    myGEM.gem()
    logging.info("Gibbs-EM finished in %.2f seconds.\n", time.time() - start)

    if is_root and not os.path.exists(args.output_directory):
        os.makedirs(args.output_directory)
    gem2csv(args.output_directory, myGEM, prefix=args.output_prefix)
    logging.info("Results are under directory %s." % args.output_directory)
    logging.info("Logging info is written to %s.", args.logging_file)
    logging.info("#" * 80 + "\n")
    if distributed:
        import torch.distributed as td
        td.destroy_process_group()
    return myGEM


if __name__ == "__main__":
    main()
