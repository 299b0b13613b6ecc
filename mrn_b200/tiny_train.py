"""Driver glue of the reference (tiny_train.py:195-294,404-460) around the CUDA path: config loading without mmcv,
the task loop (incremental_train -> test -> after_task) and the seed / directory setup.

The dataset layer (data/data_manage.py, data/dataset.py: LMDB readers, augmentation) is outside the hot path
(SURVEY.md §8f.2): `train()` takes any object with the Dataset_Manager surface the learner calls --
`init_start(opt, select_data, log, taski)`, `get_dataset(taski, memory=..., index_list=...)`, `get_batch()`,
`get_batch2()` -- plus a factory of validation loaders, so the reference's own Dataset_Manager / Val_Dataset plug in
unchanged and the tests drive the loop with in-memory synthetic data.

    python -m mrn_b200.tiny_train --config config/svtr_mrn.py      # needs the reference's data package + LMDB datasets
"""
import argparse
import os
import random
import runpy
import sys

import numpy as np
import torch


def load_config(path: str) -> argparse.Namespace:
    """mmcv.Config.fromfile + the merge of tiny_train.py:410-417: the config is a Python file defining the dicts
    `common`, `model`, `train`, `optimizer` (config/svtr_mrn.py, config/crnn_mrn.py)."""
    ns = runpy.run_path(path)
    opt = {}
    for section in ("common", "model", "train", "optimizer"):
        if section not in ns or not isinstance(ns[section], dict):
            raise ValueError("config %s has no dict `%s`" % (path, section))
        opt.update(ns[section])
    return argparse.Namespace(**opt)


def seed_everything(seed: int) -> None:
    """tiny_train.py:419-425."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def load_dict(path: str, char: dict):
    """tiny_train.py:37-53: cumulative character list in dict-file insertion order."""
    with open(os.path.join(path, "dict.txt")) as f:
        for line in f:
            ch = line.strip("\n")
            if char.get(ch) is None:
                char[ch] = 1
    return list(char.keys()), char


def train(opt, data_manager, make_valid_loader, characters_for_task, make_test_loaders=None, log=None):
    """The `il == "mrn"` branch of tiny_train.py:195-294.

    data_manager           -- the train loader object handed to MRN.incremental_train
    make_valid_loader(k)   -- validation loader object for tasks 0..k (Val_Dataset in the reference)
    characters_for_task(k) -- cumulative character list after task k (load_dict in the reference)
    make_test_loaders(k)   -- list of (images, labels) iterables, one per benchmark set (MRN.test); optional
    Returns (best_scores, ned_scores)."""
    from .il_modules.mrn import MRN
    if getattr(opt, "il", "mrn") != "mrn":
        raise NotImplementedError("mrn_b200 implements the MRN learner (il='mrn'); got %r" % (opt.il,))
    os.makedirs(f"./saved_models/{opt.exp_name}", exist_ok=True)
    learner = MRN(opt)
    best_scores, ned_scores = [], []
    for taski in range(len(opt.lan_list)):
        valid_loader = make_valid_loader(taski)
        if taski == 0 and hasattr(data_manager, "init_start"):
            data_manager.init_start(opt, getattr(opt, "select_data", None), log, taski)
        opt.character = characters_for_task(taski)
        learner.incremental_train(taski, opt.character, data_manager, valid_loader)
        if make_test_loaders is not None:
            best_scores, ned_scores = learner.test(None, make_test_loaders(taski), best_scores, ned_scores, taski)
        learner.after_task()
    if best_scores:
        print("ALL Average Incremental Accuracy: {:.2f} \n".format(sum(best_scores) / len(best_scores)))
    return best_scores, ned_scores


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="config/svtr_mrn.py")
    a = ap.parse_args(argv)
    opt = load_config(a.config)
    seed_everything(opt.manual_seed)
    if not torch.cuda.is_available():
        raise RuntimeError("mrn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    from . import dist as mdist
    _, local_rank, _ = mdist.init_from_env()     # one process per GPU under torchrun; a no-op for a single process
    torch.cuda.set_device(local_rank)
    # the input pipeline with the reference's protocol (LMDB readers, rehearsal memory, get_batch / get_batch2) and the
    # device-side AlignCollate: mrn_b200/data_manage.py
    from .data_manage import Dataset_Manager, Val_Dataset, hierarchical_dataset
    from .data import AlignCollate
    char = {}
    chars = {}

    def characters_for_task(k):
        c = None
        for data_path in opt.select_data:
            c, _ = load_dict(os.path.join(data_path, opt.lan_list[k]), char)
        chars[k] = c
        return c

    valid_datas = []

    def make_valid_loader(k):
        for valid_data in opt.valid_datas:
            valid_datas.append(os.path.join(valid_data, opt.lan_list[k]))
        return Val_Dataset(valid_datas, opt)

    def make_test_loaders(k):
        from .data_manage import _Collated, _passthrough
        collate = AlignCollate(opt, mode="test")             # device-side resize in the main process
        out = []
        for v in valid_datas:
            ds, _ = hierarchical_dataset(root=v, opt=opt, mode="test")
            out.append(_Collated(torch.utils.data.DataLoader(ds, batch_size=opt.batch_size, shuffle=True, num_workers=int(opt.workers),
                                                             collate_fn=_passthrough, pin_memory=False), collate))
        return out

    log = open(f"./saved_models/{opt.exp_name}/log_train.txt", "a") if os.path.isdir(f"./saved_models/{opt.exp_name}") else None
    train(opt, Dataset_Manager(opt), make_valid_loader, characters_for_task, make_test_loaders, log)


if __name__ == "__main__":
    main(sys.argv[1:])
