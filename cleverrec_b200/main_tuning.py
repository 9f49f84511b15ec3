# coding: utf-8
""" Grid search -- same flow as the reference main_tuning.py:16-66: `embed_size`, `reg` and `neg_ratio` are bracketed lists in the
model's .properties file, every combination (nested in that order, :41-43) builds a fresh model on the same preprocessed data and runs
`run_model()`.

What changed underneath: the reference runs the combinations one after the other, each in a new tf.Session after
tf.reset_default_graph() (:48-55).  Here a combination is one model object on its own device handle (Engine) and its own CUDA stream,
so they are independent by construction and run concurrently:
  * `tuning.workers=N` (default 1) host threads on one GPU -- every numeric call is a ctypes call into libcleverrec_b200.so (the GIL is
    released), so small models that cannot fill 148 SMs alone overlap on the device;
  * under torchrun (WORLD_SIZE = G) the grid is dealt round-robin to the ranks, one GPU each, every model running as a single-GPU
    replica (`dist.mode=replica`), and rank 0 gathers the results.
Results come back in grid order whatever the execution order was.  `sampler=numpy_stream` draws from NumPy's global stream and therefore
needs tuning.workers=1."""
import importlib
import itertools
import os
import sys
import threading
from concurrent.futures import ThreadPoolExecutor


def _list(value, cast):
    s = str(value).strip()
    if s.startswith('[') and s.endswith(']'):
        return [cast(x) for x in s[1:-1].split(',') if x.strip()]
    return [cast(s)]      # a scalar is a one-point axis (the reference requires the brackets, main_tuning.py:38-40)


def grid(configs):
    """[{'embed_size': int, 'reg': float, 'neg_ratio': int}, ...] in the reference's loop order (main_tuning.py:38-45)."""
    axes = (_list(configs['embed_size'], int), _list(configs['reg'], float), _list(configs['neg_ratio'], int))
    return [{'embed_size': e, 'reg': r, 'neg_ratio': n} for e, r, n in itertools.product(*axes)]


def _run_one(configs, combo, data, logger, device):
    import torch
    cfg = dict(configs)
    cfg.update(combo)                       # main_tuning.py:44-45 (values stay int / float: every model casts them itself)
    cfg['engine.device'] = str(device)
    cfg['dist.mode'] = 'replica'            # a tuning worker is a single-GPU replica even under torchrun
    module = 'cleverrec_b200.model.' + cfg['model_type'] + '.' + cfg['recommender']
    if importlib.util.find_spec(module) is None:
        raise Exception('Module %s not found.' % module)
    cls = getattr(importlib.import_module(module), cfg['recommender'])
    torch.cuda.set_device(device)
    with torch.cuda.stream(torch.cuda.Stream(device=device)):   # the current stream is per thread: one stream per combination
        model = cls(None, data, cfg, logger)
        try:
            best_epoch, best_metrics = model.run_model()
        finally:
            torch.cuda.current_stream().synchronize()
            model.engine.close()
    return {'params': combo, 'best_epoch': best_epoch, 'best_metrics': best_metrics}


def run_grid(configs, data, logger, workers=None, rank=None, world=None):
    """Runs every combination; returns the list of {'params', 'best_epoch', 'best_metrics'} in grid order (on every rank)."""
    combos = grid(configs)
    workers = int(configs.get('tuning.workers', 1)) if workers is None else int(workers)
    rank = int(os.environ.get('RANK', '0')) if rank is None else rank
    world = int(os.environ.get('WORLD_SIZE', '1')) if world is None else world
    device = int(configs.get('engine.device', os.environ.get('LOCAL_RANK', '0') if world > 1 else 0))
    if workers > 1 and configs.get('sampler', 'philox') == 'numpy_stream':
        raise ValueError('sampler=numpy_stream draws from the global NumPy stream: use tuning.workers=1')
    mine = list(range(rank, len(combos), world))
    lock = threading.Lock()

    def job(k):
        out = _run_one(configs, combos[k], data, logger, device)
        with lock:
            logger.info('[tuning %d/%d] %s -> best_epoch %d' % (k + 1, len(combos), combos[k], out['best_epoch']))
        return k, out

    if workers > 1:
        with ThreadPoolExecutor(max_workers=workers) as pool:
            local = list(pool.map(job, mine))
    else:
        local = [job(k) for k in mine]
    results = [None] * len(combos)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('WORLD_SIZE > 1: call torch.distributed.init_process_group before run_grid')
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        local = [kv for part in gathered for kv in part]
    for k, out in local:
        results[k] = out
    return results


def best_of(results, k_id=0):
    """The combination with the best NDCG@topk[k_id] (what one reads off the reference's log by eye)."""
    scored = [r for r in results if r and k_id in r['best_metrics']]
    return max(scored, key=lambda r: r['best_metrics'][k_id][2]) if scored else None


if __name__ == '__main__':
    from cleverrec_b200.main import load_configs
    from cleverrec_b200.utils.tools import get_logger
    root = sys.argv[1] if len(sys.argv) > 1 else '.'
    configs = load_configs(root)
    logger = get_logger(configs['log.dir'], configs['recommender'])
    logger.info('=' * 100)
    logger.info('Current model: %s' % configs['recommender'])
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl')
    if configs.get('data.preprocess', 'packaged') == 'reference':
        sys.path.insert(0, root)
        from model.RankingPreprocess import RankingPreprocess
    else:
        from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    data = RankingPreprocess(configs, logger)
    results = run_grid(configs, data, logger)
    if int(os.environ.get('RANK', '0')) == 0:
        for r in results:
            logger.info('%s best_epoch=%d %s' % (r['params'], r['best_epoch'], {k: tuple(round(x, 4) for x in v) for k, v in r['best_metrics'].items()}))
        b = best_of(results)
        if b:
            logger.info('best by NDCG@topk[0]: %s' % (b['params'],))
