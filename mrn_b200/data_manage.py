"""Input pipeline of the MRN learner: the reference's data/data_manage.py (Dataset_Manager,
Val_Dataset, IndexConcatDataset) and the LMDB side of data/dataset.py (LmdbDataset, hierarchical_dataset,
AlignCollate2) behind the same names and call protocol, so that `tiny_train.train(opt, Dataset_Manager(opt), ...)`
runs without the reference tree on sys.path.

What differs from the reference, on purpose:
  * the loaders are stepped with `next(it)` -- the reference calls `data_loader_iter.next()` (data/data_manage.py:181,
    204), which no longer exists in torch >= 1.13, so its data path does not run on this image's torch 2.11;
  * DataLoader workers only read + decode (LMDB get, PIL open, RGBA convert: host work); the bicubic resize, ToTensor
    and normalisation run on the device in the main process (mrn_b200.data.AlignCollate -> mrnb_resize_normalize_rgba,
    byte-identical to the PIL path), since CUDA cannot be used from forked workers and the host resize would bound a
    15 ms step;
  * augmentations (opt.Aug != "None") are outside the hot path and raise (config/*_mrn.py use Aug="None").

Record format (tools/create_lmdb_dataset.py:323-348): keys `num-samples`, `image-%09d`, `label-%09d` (1-based).
The `lmdb` package is not part of this image; LmdbDataset imports it lazily and raises a clear error without it.
"""
import bisect
import io
import os
from typing import List, Optional

import numpy as np
import torch
from torch.utils.data import ConcatDataset, DataLoader, Dataset, Subset


def _open_env(root):
    try:
        import lmdb
    except Exception as ex:                      # pragma: no cover - exercised only without the package
        raise RuntimeError("reading %s needs the `lmdb` package (data/dataset.py:51-58), which is not installed: %s" % (root, ex))
    env = lmdb.open(root, max_readers=32, readonly=True, lock=False, readahead=False, meminit=False)
    if not env:
        raise RuntimeError("cannot open lmdb from %s" % (root,))
    return env


class LmdbDataset(Dataset):
    """data/dataset.py:44-112: one LMDB directory; samples whose label is missing or longer than opt.batch_max_length are
    filtered out at construction; __getitem__ -> (PIL RGBA image, label str); an undecodable image becomes a blank
    imgW x imgH crop labelled "[dummy_label]"."""

    def __init__(self, root, opt, mode="train"):
        self.root, self.opt, self.mode = root, opt, mode
        self.env = _open_env(root)
        self.filtered_index_list: List[int] = []
        with self.env.begin(write=False) as txn:
            n = int(txn.get("num-samples".encode()))
            for index in range(1, n + 1):                        # lmdb keys are 1-based
                raw = txn.get(b"label-%09d" % index)
                if raw is None:
                    continue
                if len(raw.decode("utf-8")) > opt.batch_max_length:
                    continue
                self.filtered_index_list.append(index)
        self.nSamples = len(self.filtered_index_list)

    def __len__(self):
        return self.nSamples

    def __getitem__(self, index):
        from PIL import Image
        if not 0 <= index < len(self):
            raise IndexError("index range error")
        key = self.filtered_index_list[index]
        with self.env.begin(write=False) as txn:
            label = txn.get(b"label-%09d" % key).decode("utf-8")
            imgbuf = txn.get(b"image-%09d" % key)
        try:
            img = Image.open(io.BytesIO(imgbuf)).convert("RGBA")
        except (IOError, TypeError):
            img = Image.new("RGBA", (self.opt.imgW, self.opt.imgH))
            label = "[dummy_label]"
        return img, label

    # DataLoader workers re-open the environment (an lmdb handle must not cross a fork)
    def __getstate__(self):
        d = dict(self.__dict__)
        d["env"] = None
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.env = _open_env(self.root)


def hierarchical_dataset(root, opt, select_data="/", data_type="label", mode="train", dataset_cls=None):
    """data/dataset.py:15-41: every leaf directory under root whose path contains one of select_data is one LmdbDataset."""
    dataset_cls = dataset_cls or LmdbDataset
    datasets, log = [], f"dataset_root:  {root}\t dataset: {select_data}\n"
    for dirpath, dirnames, _ in os.walk(root + "/"):
        if dirnames or not any(sel in dirpath for sel in select_data):
            continue
        ds = dataset_cls(dirpath, opt, mode=mode)
        line = f"sub-directory:\t/{os.path.relpath(dirpath, root)}\t num samples: {len(ds)}"
        print(line)
        log += line + "\n"
        datasets.append(ds)
    return ConcatDataset(datasets), log


class IndexConcatDataset(ConcatDataset):
    """data/data_manage.py:272-283: ConcatDataset whose items carry the index of the member dataset they came from --
    the router's domain label: 0 = rehearsal memory of earlier tasks, 1 = current task (reference quirk 3)."""

    def __getitem__(self, idx):
        if idx < 0:
            if -idx > len(self):
                raise ValueError("absolute value of index should not exceed dataset length")
            idx = len(self) + idx
        d = bisect.bisect_right(self.cumulative_sizes, idx)
        return self.datasets[d][idx if d == 0 else idx - self.cumulative_sizes[d - 1]], d


def _passthrough(batch):
    return batch


class Dataset_Manager(object):
    """data/data_manage.py:8-217.  dataset_cls(root, opt, mode=) builds one per-language dataset (LmdbDataset unless a
    test injects another Dataset); collate(list of (image, label)) -> (image tensor, labels) defaults to the device-side
    mrn_b200.data.AlignCollate."""

    def __init__(self, opt, dataset_cls=None, collate=None):
        self.data_list = []
        self.data_loader_list = []
        self.dataloader_iter_list = []
        self.loader_kinds: List[str] = []
        self.select_data = getattr(opt, "select_data", None)
        self.opt = opt
        self.dataset_cls = dataset_cls or LmdbDataset
        self._collate = collate

    # -- collate: device-side resize + normalise in the main process ------------------------------------------------
    def collate(self, pairs):
        if self._collate is None:
            from .data import AlignCollate
            self._collate = AlignCollate(self.opt)
        return self._collate(pairs)

    # -- dataset assembly -------------------------------------------------------------------------------------------
    def create_dataset(self, data_list="/", taski=0, mode="train", repeat=True):
        """One ConcatDataset over the task's language in every data root; small datasets are repeated up to 50 k samples
        (data/data_manage.py:127-146)."""
        out = []
        for data_root in data_list:
            ds = self.dataset_cls(data_root + "/" + self.opt.lan_list[taski], self.opt, mode=mode)
            print(f"num samples: {len(ds)}")
            if 0 < len(ds) < 50000 and repeat:
                ds = ConcatDataset([ds] * int(50000 / len(ds)))
            out.append(ds)
        return ConcatDataset(out)

    def rehearsal_prev_model(self, taski):
        ds = self.create_dataset(data_list=self.select_data, taski=taski - 1, repeat=False)
        return self._loader(ds, self.opt.batch_size, shuffle=False), len(ds)

    def rehearsal_memory(self, taski, random=False, total_num=2000, index_array=None, repeat=False):
        """Subsets of every earlier task's dataset: index_array[i] (or a fresh random draw) picks task i's samples."""
        per_task = int(total_num / taski)
        print("memory size is {}\n".format(per_task))
        parts, used = [], []
        for i in range(taski):
            ds = self.create_dataset(data_list=self.select_data, taski=i, repeat=repeat)
            idx = np.random.choice(range(len(ds)), per_task, replace=repeat) if random else np.asarray(index_array[i])
            parts.append(Subset(ds, idx.tolist()))
            used.append(idx)
        return ConcatDataset(parts), used

    def _loader(self, dataset, batch_size, shuffle=True):
        return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=int(getattr(self.opt, "workers", 0)),
                          collate_fn=_passthrough, pin_memory=False, drop_last=False)

    def create_dataloader(self, dataset, batch_size=None):
        dl = self._loader(dataset, self.opt.batch_size if batch_size is None else batch_size)
        self.data_loader_list.append(dl)
        self.dataloader_iter_list.append(iter(dl))
        self.loader_kinds.append("plain")

    def create_dataloader_mix(self, dataset, batch_size=None):
        dl = self._loader(dataset, self.opt.batch_size if batch_size is None else batch_size)
        self.data_loader_list.append(dl)
        self.dataloader_iter_list.append(iter(dl))
        self.loader_kinds.append("mix")

    def get_dataset(self, taski, memory="random", index_list=None):
        """data/data_manage.py:16-61: (re)builds the loaders of a stage.  memory=None: the task's dataset alone (stage 0);
        memory set and il == "mrn": IndexConcatDataset([rehearsal memory, memory_num / taski samples of the task]) for the
        router stage; "large" / "total" / "test_ch" as in the reference; any other value: two half-batch loaders."""
        self.data_loader_list, self.dataloader_iter_list, self.loader_kinds = [], [], []
        memory_num = self.opt.memory_num if memory is not None else 0
        dataset = self.create_dataset(data_list=self.select_data, taski=taski)
        if memory is not None and getattr(self.opt, "il", "mrn") == "mrn":
            cur = np.random.choice(range(len(dataset)), int(memory_num / taski), replace=False)
            mem, index_list = self.rehearsal_memory(taski, random=False, total_num=memory_num, index_array=index_list)
            self.create_dataloader_mix(IndexConcatDataset([mem, Subset(dataset, cur.tolist())]), self.opt.batch_size)
        elif memory == "test_ch":
            mem, index_list = self.rehearsal_memory(taski, random=False, total_num=memory_num, index_array=index_list, repeat=True)
            self.create_dataloader_mix(IndexConcatDataset([mem, dataset]), self.opt.batch_size)
        elif memory == "large":
            cur = np.random.choice(range(len(dataset)), memory_num, replace=False)
            mem, index_list = self.rehearsal_memory(taski, random=False, total_num=memory_num * taski, index_array=index_list)
            self.create_dataloader_mix(IndexConcatDataset([mem, Subset(dataset, cur.tolist())]), self.opt.batch_size)
        elif memory == "total":
            parts = [dataset] + [self.create_dataset(data_list=self.select_data, taski=i) for i in range(taski)]
            self.create_dataloader_mix(IndexConcatDataset(parts), self.opt.batch_size)
        elif memory is not None:
            mem, index_list = self.rehearsal_memory(taski, random=False, total_num=memory_num, index_array=index_list)
            self.create_dataloader(mem, self.opt.batch_size // 2)
            self.create_dataloader(dataset, self.opt.batch_size // 2)
        else:
            self.create_dataloader(dataset)
        return index_list

    def init_start(self, opt, select_data, log, taski):
        self.opt, self.select_data = opt, select_data
        line = "-" * 80 + "\n" + f"select_data: {select_data}\n"
        print(line)
        if log is not None:
            log.write(line)
        self.get_dataset(taski, memory=None)

    # -- batches ----------------------------------------------------------------------------------------------------
    def _next(self, i):
        """next batch of loader i, restarting an exhausted loader (the reference's try / except StopIteration with the
        removed iterator method `.next()` replaced by the builtin)."""
        try:
            return next(self.dataloader_iter_list[i])
        except StopIteration:
            self.dataloader_iter_list[i] = iter(self.data_loader_list[i])
            return next(self.dataloader_iter_list[i])

    def get_batch(self):
        """-> (images [B,4,imgH,imgW], labels): one batch from every loader, concatenated (data/data_manage.py:198-217)."""
        images, labels = [], []
        for i in range(len(self.dataloader_iter_list)):
            img, lab = self.collate(self._next(i))
            images.append(img)
            labels += list(lab)
        return torch.cat(images, 0), labels

    def get_batch2(self):
        """-> (images, labels, [domain index tensor per loader]) for the router stage (data/data_manage.py:174-196)."""
        images, labels, index = [], [], []
        for i in range(len(self.dataloader_iter_list)):
            pairs, dom = zip(*self._next(i))                     # IndexConcatDataset items: ((image, label), dataset_idx)
            img, lab = self.collate(list(pairs))
            images.append(img)
            labels += list(lab)
            index.append(torch.tensor(dom, dtype=torch.long))
        return torch.cat(images, 0), labels, index


class Val_Dataset(object):
    """data/data_manage.py:219-269: validation loaders over the benchmark directories seen so far."""

    def __init__(self, val_datas, opt, dataset_cls=None, collate=None):
        self.val_datas, self.opt = val_datas, opt
        self.current_data = val_datas[-1]
        self.dataset_cls = dataset_cls
        self._collate = collate

    def _collate_fn(self):
        if self._collate is None:
            from .data import AlignCollate
            self._collate = AlignCollate(self.opt, mode="test")
        return self._collate

    def _loader(self, dataset):
        # collate in the main process (device-side resize); workers would need CUDA otherwise
        return _Collated(DataLoader(dataset, batch_size=self.opt.batch_size, shuffle=True,
                                    num_workers=int(getattr(self.opt, "workers", 0)), collate_fn=_passthrough, pin_memory=False),
                         self._collate_fn())

    def create_dataset(self, val_data=None):
        ds, _ = hierarchical_dataset(root=val_data or self.current_data, opt=self.opt, mode="test", dataset_cls=self.dataset_cls)
        print("-" * 80)
        return self._loader(ds)

    def create_list_dataset(self, valid_datas=None):
        parts = []
        for val_data in (valid_datas or self.val_datas):
            ds, log = hierarchical_dataset(root=val_data, opt=self.opt, mode="test", dataset_cls=self.dataset_cls)
            if len(ds) > 700:                                   # at most 700 samples per benchmark set while training
                ds = Subset(ds, np.random.choice(range(len(ds)), 700, replace=False).tolist())
            parts.append(ds)
            print(log + "-" * 80)
        return self._loader(ConcatDataset(parts))


class _Collated:
    """Iterable of collated batches over a passthrough DataLoader (len() = number of batches, as validation() expects)."""

    def __init__(self, loader, collate):
        self.loader, self.collate = loader, collate

    def __len__(self):
        return len(self.loader)

    def __iter__(self):
        for pairs in self.loader:
            yield self.collate(pairs)
