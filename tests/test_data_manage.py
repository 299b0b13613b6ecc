"""Host-side input pipeline: mrn_b200.data_manage mirrors data/data_manage.py + the LMDB side of
data/dataset.py.  No GPU: a CPU collate stands in for the device-side AlignCollate, and a dict-backed fake `lmdb`
module stands in for the package (absent from this image) to pin the record format and the filtering rules."""
import argparse
import io
import sys
import types

import numpy as np
import pytest
import torch
from PIL import Image
from torch.utils.data import Dataset

from mrn_b200 import data_manage as dm


def make_opt(**kw):
    d = dict(lan_list=["Chinese", "Latin", "Bangla"], batch_size=8, workers=0, memory_num=12, il="mrn", imgH=32, imgW=256,
             batch_max_length=25, select_data=["rootA", "rootB"], Aug="None")
    d.update(kw)
    return argparse.Namespace(**d)


class FakeLang(Dataset):
    """20 + 3 * task crops per (root, language); label encodes where the sample came from."""
    def __init__(self, root, opt, mode="train"):
        self.root = root
        lang = root.rstrip("/").rsplit("/", 1)[-1]
        self.n = 20 + 3 * opt.lan_list.index(lang)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return Image.new("RGBA", (40 + i, 20), (i, 0, 0, 255)), "%s#%d" % (self.root, i)


def cpu_collate(pairs):
    images, labels = zip(*pairs)
    t = [torch.from_numpy(np.asarray(im.resize((256, 32)), dtype=np.float32)).permute(2, 0, 1) for im in images]
    return torch.stack(t, 0), labels


def test_get_batch_restarts_exhausted_loaders():
    opt = make_opt()
    m = dm.Dataset_Manager(opt, dataset_cls=FakeLang, collate=cpu_collate)
    m.init_start(opt, opt.select_data, None, 0)
    n_total = len(m.data_loader_list[0].dataset)
    assert n_total >= 50000                         # 2 roots x 20 samples, each repeated up to 50 k (data_manage.py:137-141)
    img, labels = m.get_batch()
    assert img.shape == (8, 4, 32, 256) and len(labels) == 8
    # exhaust a tiny loader and keep going: the reference's iterator `.next()` is gone in torch >= 1.13; next(it) + restart
    m.data_loader_list, m.dataloader_iter_list = [], []
    m.create_dataloader(FakeLang("rootA/Chinese", opt), batch_size=8)
    seen = [m.get_batch()[0].shape[0] for _ in range(7)]     # 20 samples -> 8, 8, 4, then restarted
    assert seen == [8, 8, 4, 8, 8, 4, 8]


def test_router_stage_batches_carry_domain_indices():
    opt = make_opt()
    m = dm.Dataset_Manager(opt, dataset_cls=FakeLang, collate=cpu_collate)
    np.random.seed(0)
    index_list = [np.arange(6), np.arange(3, 9)]            # rehearsal memory: 6 samples of each earlier task (memory_num / taski)
    used = m.get_dataset(2, memory="random", index_list=index_list)
    assert [u.tolist() for u in used] == [ix.tolist() for ix in index_list]
    ds = m.data_loader_list[0].dataset
    assert isinstance(ds, dm.IndexConcatDataset) and len(ds) == 12 + 6   # memory 2 x 6 + memory_num / taski of the current task
    img, labels, index = m.get_batch2()
    assert img.shape[0] == 8 and len(labels) == 8 and len(index) == 1 and index[0].dtype == torch.long
    dom = index[0].tolist()
    assert set(dom) <= {0, 1}
    for lab, d in zip(labels, dom):                          # 0 = rehearsal memory (earlier languages), 1 = current task
        assert ("Bangla" in lab) == (d == 1)
    # the learner flattens the per-loader index list like the reference's LongTensor(indexs).squeeze()
    from mrn_b200.il_modules.mrn import _domain_ids
    assert _domain_ids(index).tolist() == dom
    _, n_prev = m.rehearsal_prev_model(2)
    assert n_prev == 2 * (20 + 3)                            # task 1 dataset, not repeated


def test_lmdb_dataset_record_format_and_filters(monkeypatch):
    def png(w, h):
        b = io.BytesIO()
        Image.new("RGB", (w, h), (10, 20, 30)).save(b, format="PNG")
        return b.getvalue()
    store = {b"num-samples": b"4", b"label-000000001": "ok".encode(), b"image-000000001": png(50, 20),
             b"label-000000002": ("x" * 26).encode(), b"image-000000002": png(50, 20),      # longer than batch_max_length
             b"image-000000003": png(50, 20),                                                # label missing
             b"label-000000004": "汉字".encode("utf-8"), b"image-000000004": b"not an image"}

    class Txn:
        def __enter__(self): return self
        def __exit__(self, *a): return False
        def get(self, k): return store.get(k)

    class Env:
        def begin(self, write=False): return Txn()

    fake = types.ModuleType("lmdb")
    fake.open = lambda root, **kw: Env()
    monkeypatch.setitem(sys.modules, "lmdb", fake)
    ds = dm.LmdbDataset("some/dir", make_opt())
    assert len(ds) == 2 and ds.filtered_index_list == [1, 4]
    img, label = ds[0]
    assert img.mode == "RGBA" and img.size == (50, 20) and label == "ok"
    img, label = ds[1]                                        # undecodable image -> blank crop + dummy label
    assert img.size == (256, 32) and label == "[dummy_label]"
    with pytest.raises(IndexError):
        ds[2]


def test_val_dataset_lists_are_subsampled(tmp_path):
    for name in ("test_2017/Chinese", "test_2017/Latin"):
        (tmp_path / name).mkdir(parents=True)

    class Big(FakeLang):
        def __len__(self):
            return 900

        def __getitem__(self, i):
            return super().__getitem__(i % 40)
    opt = make_opt(batch_size=100)
    v = dm.Val_Dataset([str(tmp_path / "test_2017" / "Chinese"), str(tmp_path / "test_2017" / "Latin")], opt, dataset_cls=Big,
                       collate=cpu_collate)
    np.random.seed(1)
    loader = v.create_list_dataset()
    assert len(loader) == 14                                  # 2 x 700 samples in batches of 100
    img, labels = next(iter(loader))
    assert img.shape == (100, 4, 32, 256) and len(labels) == 100
    single = v.create_dataset()
    assert len(single) == 9
