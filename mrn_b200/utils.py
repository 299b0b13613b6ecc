"""Label codec and small helpers mirroring tools/utils.py of the reference (host-side logic)."""
import torch


class CTCLabelConverter(object):
    """Same vocabulary contract as tools/utils.py:10-76: 0 = [CTCblank], 1 = [PAD], 2 = [UNK], 3 = ' ', chars from 4.
    encode() builds the [B, batch_max_length] int64 index tensor padded with [PAD]; decode() accepts either the raw
    per-frame arg-max (reference behaviour: collapse repeats, drop blank) or the compact ids produced on the device
    by mrnb_greedy_decode (decode_compact)."""

    def __init__(self, character, device=None):
        list_special_token = ["[PAD]", "[UNK]", " "]
        dict_character = list_special_token + list(character)
        self.dict = {char: i + 1 for i, char in enumerate(dict_character)}
        self.character = ["[CTCblank]"] + dict_character
        self.device = device

    def encode(self, word_string, batch_max_length=25, device=None):
        """tools/utils.py:35-60.  device="cpu" keeps the tensors on the host (for the prefetcher)."""
        word_length = [len(word) for word in word_string]
        word_index = torch.full((len(word_string), batch_max_length), self.dict["[PAD]"], dtype=torch.long)
        for i, word in enumerate(word_string):
            idx = [self.dict.get(ch, self.dict["[UNK]"]) for ch in word]
            word_index[i, :len(idx)] = torch.tensor(idx, dtype=torch.long)
        lens = torch.tensor(word_length, dtype=torch.int32)
        target = self.device if device is None else (None if str(device) == "cpu" else device)
        if target is not None:
            word_index, lens = word_index.to(target, non_blocking=True), lens.to(target, non_blocking=True)
        return word_index, lens

    def decode(self, word_index, word_length):
        word_index = word_index.cpu().tolist() if isinstance(word_index, torch.Tensor) else word_index
        out = []
        for idx, length in enumerate(word_length):
            row = word_index[idx]
            chars = [self.character[row[i]] for i in range(int(length))
                     if row[i] != 0 and not (i > 0 and row[i - 1] == row[i])]
            out.append("".join(chars))
        return out

    def decode_compact(self, ids, lens):
        """ids [B,T] (collapsed on the device, -1 padded), lens [B]: ONE device->host copy, then a join."""
        ids, lens = ids.cpu().tolist(), lens.cpu().tolist()
        return ["".join(self.character[c] for c in row[:n]) for row, n in zip(ids, lens)]


class Averager(object):
    """tools/utils.py Averager: running mean of scalar tensors (kept on the device until val())."""

    def __init__(self):
        self.reset()

    def add(self, v):
        v = v.detach().float().sum() if isinstance(v, torch.Tensor) else torch.tensor(float(v))
        self.sum = v if self.sum is None else self.sum + v
        self.n_count += 1

    def reset(self):
        self.n_count = 0
        self.sum = None

    def val(self):
        if self.n_count == 0:
            return 0.0
        return float(self.sum) / float(self.n_count)


class DevicePrefetcher:
    """Double-buffered host -> device staging on a side stream: batch k+1 is copied (from pinned memory, non-blocking)
    while the kernels of batch k run, and the compute stream only waits on the copy's event.  Replaces the blocking
    `image_tensors.to(device)` at the top of the reference loops (il_modules/mrn.py:236-238,331-335)."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.pending = None

    def submit(self, tensors):
        """Start copying a tuple of host tensors; returns nothing.  Call take() to obtain the device tensors."""
        with torch.cuda.stream(self.stream):
            dev = tuple(t.to(self.device, non_blocking=True) if isinstance(t, torch.Tensor) else t for t in tensors)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.pending = (dev, ev)

    def take(self):
        dev, ev = self.pending
        self.pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in dev:
            if isinstance(t, torch.Tensor):
                t.record_stream(cur)
        return dev
