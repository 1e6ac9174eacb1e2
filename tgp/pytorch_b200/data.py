"""Host -> device minibatch staging for datasets that live in host memory.

The reference feeds the hot path from a `torch.utils.data.DataLoader` (`code/dsp/data/data.py:27-61`): every step
gathers `X[idx], Y[idx]` on the host and copies them to the device synchronously, so the GPU idles during the gather
and the copy.  `PinnedMinibatchStager` keeps the same indexing (the caller supplies the index tensor of every step, so
"identical minibatch indexing" is the caller's) and overlaps it with compute:

    stager = PinnedMinibatchStager(X_host, Y_host, rows, device)
    stager.stage(idx_0)
    for s in range(steps):
        xb, yb = stager.get()                # device tensors of step s (the compute stream waits for the copy)
        loss = -model.ELBO(xb, yb)[0]; loss.backward()          # kernels of step s enqueued (asynchronous)
        stager.stage(idx_{s+1})              # host gather + H2D of step s+1 while they run
        optimizer.step(); loss.item() ...

Two pinned host buffers and two device buffers alternate.  The H2D copies run on a private copy stream; events order
them against the compute stream in both directions (a device buffer is not overwritten before the step that read it has
finished; a step does not start before its copy has landed).  No CPU fallback: the device must be CUDA.
"""
import torch


class PinnedMinibatchStager:
    def __init__(self, X, Y, rows, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('PinnedMinibatchStager stages into CUDA memory; got device %s' % device)
        if X.is_cuda or Y.is_cuda:
            raise ValueError('the dataset is expected in host memory (use plain indexing for device-resident data)')
        self.X, self.Y = X, Y.view(X.shape[0], -1)
        self.rows, self.device = int(rows), device
        d, dy = X.shape[1], self.Y.shape[1]
        self._xh = [torch.empty(rows, d, dtype=X.dtype).pin_memory() for _ in range(2)]
        self._yh = [torch.empty(rows, dy, dtype=Y.dtype).pin_memory() for _ in range(2)]
        self._xd = [torch.empty(rows, d, dtype=X.dtype, device=device) for _ in range(2)]
        self._yd = [torch.empty(rows, dy, dtype=Y.dtype, device=device) for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=device)
        self._landed = [torch.cuda.Event() for _ in range(2)]       # H2D of the slot finished
        self._consumed = [None, None]                               # compute work that read the slot was enqueued
        self._n = [0, 0]
        self._pending_release = None
        self._staged = []                                           # slots staged and not yet handed out (FIFO)
        self._next = 0
        self.bytes_per_step = rows * (d * X.element_size() + dy * Y.element_size())

    def stage(self, idx):
        """Gather rows `idx` on the host into pinned memory and start their copy to the device."""
        if len(self._staged) == 2:
            raise RuntimeError('both staging slots are in flight: call get() before staging a third minibatch')
        k = self._next
        self._next ^= 1
        n = int(idx.numel())
        if n > self.rows:
            raise ValueError('minibatch of %d rows exceeds the stager capacity %d' % (n, self.rows))
        self._landed[k].synchronize()                               # the previous copy out of this pinned buffer is done
        torch.index_select(self.X, 0, idx, out=self._xh[k][:n])
        torch.index_select(self.Y, 0, idx, out=self._yh[k][:n])
        with torch.cuda.stream(self._copy_stream):
            if self._consumed[k] is not None:
                self._copy_stream.wait_event(self._consumed[k])     # the step that used this device buffer has run
            self._xd[k][:n].copy_(self._xh[k][:n], non_blocking=True)
            self._yd[k][:n].copy_(self._yh[k][:n], non_blocking=True)
            self._landed[k].record(self._copy_stream)
        self._n[k] = n
        self._staged.append(k)

    def get(self):
        """Device tensors (X_batch, Y_batch) of the oldest staged minibatch, valid on the current stream."""
        if not self._staged:
            raise RuntimeError('no minibatch staged')
        self.release()                                              # the previous step's kernels are enqueued by now
        k = self._staged.pop(0)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._landed[k])
        if self._consumed[k] is None:
            self._consumed[k] = torch.cuda.Event()
        self._pending_release = k
        return self._xd[k][:self._n[k]], self._yd[k][:self._n[k]]

    def release(self):
        """Mark the minibatch handed out by the last get() as consumed (its device buffers may be overwritten once the work
        enqueued so far has run).  get() does this for the previous minibatch; call it explicitly only to free a slot
        earlier."""
        k = self._pending_release
        if k is not None:
            self._consumed[k].record(torch.cuda.current_stream(self.device))
            self._pending_release = None


def load_split_csv(csv_path, split_pickle, seed, md5sum=None, sep=',', label_index=-1, pin=False):
    """Reader of the reference's large regression sets (airline: code/dsp/data/regression_datasets.py:95-192; the same on-disk
    format as its UCI sets): a header-less CSV whose `label_index` column is the target, plus `splits_idx_<name>.pkl` =
    {'seed_<k>': {'train': idx, 'test': idx}}.  Returns (X_tr, Y_tr, X_te, Y_te, Y_std) as FP64 torch tensors standardised with
    the TRAINING statistics exactly as `standard_normalization` does (data.py:260-299: numpy mean / std (ddof 0) + 1e-15),
    optionally in pinned host memory — the form `PinnedMinibatchStager` feeds from."""
    import hashlib
    import pickle
    import numpy as np
    import pandas as pd
    if md5sum is not None:
        h = hashlib.md5()
        with open(csv_path, 'rb') as fh:
            for chunk in iter(lambda: fh.read(1 << 20), b''):
                h.update(chunk)
        if h.hexdigest() != md5sum:
            raise ValueError('Dataset %s is corrupted or has not been downloaded (md5 mismatch)' % csv_path)
    data = pd.read_csv(csv_path, sep=sep, header=None).to_numpy(dtype=np.float64)
    with open(split_pickle, 'rb') as fh:
        split = pickle.load(fh)['seed_%d' % seed]
    tr, te = np.asarray(split['train']), np.asarray(split['test'])
    cols = np.ones(data.shape[1], dtype=bool)
    cols[label_index] = False
    X_tr, X_te = data[tr][:, cols], data[te][:, cols]
    Y_tr, Y_te = data[tr][:, label_index].reshape(-1, 1), data[te][:, label_index].reshape(-1, 1)
    eps = 1e-15
    x_mean, x_std = X_tr.mean(0), X_tr.std(0) + eps
    y_mean, y_std = Y_tr.mean(0), Y_tr.std(0) + eps
    out = [torch.tensor((X_tr - x_mean) / x_std), torch.tensor((Y_tr - y_mean) / y_std),
           torch.tensor((X_te - x_mean) / x_std), torch.tensor((Y_te - y_mean) / y_std)]
    if pin and torch.cuda.is_available():
        out = [t.pin_memory() for t in out]
    return out[0], out[1], out[2], out[3], torch.tensor(y_std)
