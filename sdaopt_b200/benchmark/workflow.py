"""`sda_bench` / `sdaopt_bench` command line (reference: sdaopt/benchmark/workflow.py:28-90,
setup.py:41):  python -m sdaopt_b200.benchmark.workflow --nb-runs 100 --output-folder DATA
Under torchrun the jobs are sharded over the ranks; rank 0 writes results.csv."""
import argparse
import logging
import os
import sys

DEFAULT_NB_RUNS = 100                       # workflow.py:15
DEFAULT_OUTPUT_FOLDER = os.path.join(os.getcwd(), 'DATA')   # workflow.py:16


def run_all_bench(argv=None):
    parser = argparse.ArgumentParser(description='Running SDA benchmark on the GPU')
    parser.add_argument('--nb-runs', dest='nb_runs', action='store', type=int,
                        default=DEFAULT_NB_RUNS, help='Number of runs for a given function to test')
    parser.add_argument('--output-folder', dest='output_folder', action='store',
                        default=DEFAULT_OUTPUT_FOLDER, help='Folder where data file for optimization results are stored')
    parser.add_argument('--functions', nargs='*', default=None, help='subset of the catalogue')
    parser.add_argument('--max-calls', type=int, default=int(1e6))
    parser.add_argument('--gpus', type=int, default=1,
                        help='shard the jobs over this many GPUs of the node (one process per GPU)')
    args = parser.parse_args(argv)
    if args.gpus > 1 and 'WORLD_SIZE' not in os.environ:
        # re-launch under torchrun: one rank per GPU, jobs dealt by expected cost (sharding.py)
        import subprocess
        rest = [a for a in (sys.argv[1:] if argv is None else list(argv))]
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', os.environ.get('MASTER_PORT', '29533'),
               '-m', 'sdaopt_b200.benchmark.workflow'] + rest
        return subprocess.call(cmd)
    root = logging.getLogger()
    root.setLevel(logging.INFO)
    ch = logging.StreamHandler(sys.stdout)
    ch.setFormatter(logging.Formatter('%(asctime)s - %(levelname)s - %(message)s'))   # workflow.py:50-56
    root.addHandler(ch)

    from .bench import Benchmarker
    from .benchstore import report
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    Benchmarker(args.nb_runs, args.output_folder, names=args.functions, max_calls=args.max_calls,
                rank=rank, world=world, device=local_rank).run()
    if world > 1:
        dist.barrier()
    if rank == 0:
        path = os.path.join(args.output_folder, 'results.csv')
        report(args.output_folder, kind='csv', path=path)
        logging.getLogger(__name__).info('Benchmarks done: %s', path)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    run_all_bench()
