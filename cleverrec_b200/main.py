# coding: utf-8
" CleverRec on B200: same entry flow as the reference main.py:16-55 (configs -> data -> model by name -> run_model). "
import configparser as cp
import importlib
import os
import sys


def load_configs(root='.', overrides=None):
    conf = cp.ConfigParser()
    conf.read(os.path.join(root, 'CleverRec.properties'), encoding='utf-8')
    configs = dict(conf.items('default'))
    recommender = (overrides or {}).get('recommender', configs['recommender'])
    conf.read(os.path.join(root, configs.get('config_dir', './conf'), recommender + '.properties'), encoding='utf-8')
    configs.update(dict(conf.items('parameters')))
    configs.update({k: str(v) for k, v in (overrides or {}).items()})
    return configs


def run(configs, data, logger):
    module = 'cleverrec_b200.model.' + configs['model_type'] + '.' + configs['recommender']
    if importlib.util.find_spec(module) is None:
        raise Exception('Module %s not found.' % module)
    myclass = getattr(importlib.import_module(module), configs['recommender'])
    # the `sess` slot carries the engine when the data object already owns one (data.preprocess=device), else None
    model = myclass(getattr(data, 'engine', None), data, configs, logger)
    return model.run_model()


if __name__ == '__main__':
    from cleverrec_b200.utils.tools import get_logger
    root = sys.argv[1] if len(sys.argv) > 1 else '.'
    configs = load_configs(root)
    logger = get_logger(configs['log.dir'], configs['recommender'])
    logger.info('=' * 100)
    logger.info('Current model: %s' % configs['recommender'])
    # data: the packaged mirror of model/RankingPreprocess.py (same split and evaluation negatives bit for bit under the same
    # np.random.seed -- tests/test_preprocess.py); `data.preprocess=reference` uses the reference checkout's own class instead
    if configs.get('data.preprocess', 'packaged') == 'reference':
        sys.path.insert(0, root)
        from model.RankingPreprocess import RankingPreprocess
    elif configs.get('data.preprocess', 'packaged') == 'device':   # filter / re-index / split / evaluation negatives on the GPU
        from cleverrec_b200.model.RankingPreprocess import DeviceRankingPreprocess as RankingPreprocess
    else:
        from cleverrec_b200.model.RankingPreprocess import RankingPreprocess
    data = RankingPreprocess(configs, logger)
    run(configs, data, logger)
