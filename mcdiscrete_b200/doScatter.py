"""Command line with the reference's options (MC/doScatter.py:121-252):

    python -m mcdiscrete_b200 -i input.file [-u] [-o store] [--format auto|hdf5|parquet] [--batch N] [--gpus all]

Reads the same input file, builds the tables on the GPU (or reads `../tables/integration/*.table` when
`gen_tables: no`, probTables.py:513-535), runs num_el electrons in batches and writes the 'input' / 'els' /
'scats' store (mcdiscrete_b200/store.py).  With WORLD_SIZE > 1 (torchrun) every rank simulates its id range and
writes `<store>.rankK`; histograms and counters are all-reduced and written by rank 0 as `<store>.stats.json`."""
import getopt
import json
import os
import sys
import time


def dealWithInput(argv):
    """doScatter.py:121-151: -i <file>, -u (parsed, unused), -h."""
    inputFile, use_units, out, fmt, batch = '../inputFile/input.file', False, None, 'auto', 2_000_000
    try:
        opts, _ = getopt.getopt(argv, 'uhi:o:', ['inputFile=', 'format=', 'batch=', 'output='])
    except getopt.GetoptError:
        print('doScatter.py -u -i <input file>')
        sys.exit(2)
    for opt, arg in opts:
        if opt == '-h':
            print('doScatter.py -u -i <input file> [-o <store>] [--format auto|hdf5|parquet] [--batch N]')
            sys.exit()
        elif opt == '-u':
            use_units = True
        elif opt in ('-i', '--inputFile'):
            inputFile = arg
        elif opt in ('-o', '--output'):
            out = arg
        elif opt == '--format':
            fmt = arg
        elif opt == '--batch':
            batch = int(arg)
    return use_units, inputFile, out, fmt, batch


def main(argv=None):
    from .fileTools import readInput
    from .engine import Engine
    from .probTables import genTables
    from .store import ResultStore, store_name, have_pytables
    from .distributed import init_from_env, rank_range, allreduce_stats, bind_to_gpu
    _, inputFile, out, fmt, batch = dealWithInput(sys.argv[1:] if argv is None else argv)
    inputPar = readInput(inputFile)
    rank, local, world = init_from_env()
    if world > 1:
        bind_to_gpu(local)                 # pinned staging buffers on the NUMA node of this rank's GPU
    if rank == 0:
        print(' \n - Number of incident electrons: %s' % inputPar['num_el'])
        print(' \n - Material is: %s' % inputPar['material'])
        print(' \n - Scattering mode is: %s ' % inputPar['mode'])
        print(' \n - Elastic scattering mode is: %s \n' % inputPar.get('elastic', 'Ruth_vanilla'))
    eng = Engine(inputPar, device=local)
    t0 = time.time()
    genTables(inputPar, engine=eng)
    t_tab = time.time() - t0
    use_h5 = fmt in ('hdf5', 'auto')          # PyTables when importable, else the self-contained writer (h5lite.py)
    path = out or store_name(inputPar, ext='.h5' if use_h5 else '.parquet')
    if world > 1:
        path += '.rank%d' % rank
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    lo, n = rank_range(int(inputPar['num_el']), rank, world)
    names = eng.channel_names
    t0 = time.time()
    events = 0
    import torch
    # every rank joins the one all-reduce of the run, also a rank whose id range is empty (num_el < WORLD_SIZE)
    total = torch.zeros_like(eng.stats_tensor())
    with ResultStore(path, inputPar, fmt) as st:
        if st.fmt == 'h5lite' and rank == 0:
            print(' ! PyTables is not installed: writing the HDF5 store with the self-contained writer '
                  '(mcdiscrete_b200/h5lite.py).  Its files follow the pandas format="table" layout but have not been '
                  'opened by libhdf5 / pandas on this machine; use --format parquet if in doubt.', file=sys.stderr)
        for first in range(lo, lo + n, batch):
            cnt = min(batch, lo + n - first)
            res = eng.run(count=cnt, first_id=first, pinned=True, ordered='columns')
            events += res.counters['events']
            st.append_result(res, names, double_record=bool(inputPar.get('reference_double_record', False)))
            total += eng.stats_tensor()
    allreduce_stats(total)
    dt = time.time() - t0
    if rank == 0:
        from . import _lib
        h = eng.split_hist(total[:_lib.HIST_TOTAL].cpu().numpy())
        stats = {'electrons': int(inputPar['num_el']), 'tables_s': t_tab, 'run_s': dt, 'ranks': world,
                 'histograms': {k: {'counts': v[0].tolist(), 'edges': v[1].tolist()} for k, v in h.items()}}
        json.dump(stats, open(path.replace('.rank0', '') + '.stats.json', 'w'))
        print(' BSE data had been written to ', path)
        print(' %d electrons, %.3g events on this rank in %.2f s (tables %.2f s)' % (n, events, dt, t_tab))
    eng.close()


if __name__ == '__main__':
    main()
