"""BaseWriter - mirror of gpuSolve/IO/writers/basewriter.py:6-179, rebuilt around pinned memory and the copy engine.

The reference keeps every received frame on the GPU (``tf.identity``) and moves a whole chunk to the host
with one blocking ``tf.stack(...).numpy()`` when the chunk is full (basewriter.py:62-99), then writes it -
the time loop stalls for the transfer and for the disk.  Here ``add_solution`` never blocks the stepping
stream:

  * the frame is snapshotted device-to-device into one of two staging buffers ON THE CALLER'S STREAM (so the
    solver may overwrite its own buffer immediately, the ``tf.identity`` guarantee),
  * a dedicated copy stream moves the snapshot into a PINNED host chunk buffer (DMA engine, overlaps the
    next time steps),
  * when a chunk is complete (``every_N`` frames, or ``max_chunk_mb``) a background thread waits for its last
    copy, hands the chunk to the format-specific ``_write_chunk`` and recycles the pinned buffer; the time
    loop carries on filling the second one.

Same public surface and cfg keys as the reference (``fname``, ``every_N``, ``max_chunk_mb``); derived classes
still override ``_write_chunk`` / ``_aggregate`` / ``_final_path``.  Frames may also be NumPy arrays or CPU
tensors (the reference's own round-trip test feeds NumPy rows): they are copied straight into the chunk.
"""
import os
import queue
import threading

import numpy as np
import torch

_PINNED_CAP_BYTES = 256 * 1024 * 1024      # one pinned chunk buffer never exceeds this (a logical chunk may span several)


class BaseWriter:
    _MB = 1024 * 1024

    def __init__(self, config=None):
        self._fname = 'out'
        self._every_N = None
        self._max_chunk_mb = 500
        self._buffer = []             # frames of the current (un-flushed) chunk: views into the pinned chunk buffer
        self._bw_chunk_files = []
        self._counter = 0
        self._chunk_index = 0
        self._chunk_bytes = 0
        if config is not None:
            for attribute in ['_fname', '_every_N', '_max_chunk_mb']:
                if attribute[1:] in config.keys():
                    setattr(self, attribute, config[attribute[1:]])
        # pipeline state (created on the first frame)
        self._bw_frame_shape = None
        self._bw_dtype = None
        self._bw_host = None          # two pinned (or pageable, without CUDA) chunk buffers [cap, *frame]
        self._bw_cap = 0
        self._bw_cur = 0              # buffer being filled
        self._bw_fill = 0             # frames in it
        self._bw_free = None          # semaphore: buffers the writer thread has released
        self._bw_queue = None
        self._bw_thread = None
        self._bw_error = None
        self._bw_stage = None         # two device staging buffers
        self._bw_stage_done = None    # events: D2H of the staging buffer finished
        self._bw_stage_idx = 0
        self._bw_copy_stream = None
        self._bw_last_copy = None
        self._bw_open_chunk = False   # part of the current chunk has already gone to the writer thread
        self._bw_pieces = 0           # pinned buffers handed to the writer thread so far

    # temporary chunk files written so far.  Reading it waits for the background writer, so that - as with the
    # reference's synchronous flush (basewriter.py:77-93) - every chunk flushed so far is on disk when it is looked at.
    @property
    def _chunk_files(self):
        if self._bw_queue is not None:
            self._bw_queue.join()
        return self._bw_chunk_files

    @_chunk_files.setter
    def _chunk_files(self, value):
        self._bw_chunk_files = value

    # ---- reference API --------------------------------------------------------------------------
    def set_fname(self, fname):
        self._fname = fname

    def set_every_N(self, every_N):
        self._every_N = every_N

    def set_max_chunk_mb(self, max_chunk_mb):
        self._max_chunk_mb = max_chunk_mb

    def counter(self):
        return self._counter

    def nb_chunks(self):
        return self._chunk_index

    def add_solution(self, data):
        """Receive one solution (CUDA / CPU torch tensor or ndarray): snapshot now, transfer and write in the
        background.  The caller may modify ``data`` as soon as this returns."""
        self._raise_pending()
        t = data if isinstance(data, torch.Tensor) else None
        if t is None:
            data = np.asarray(data)
        shape = tuple(t.shape) if t is not None else tuple(data.shape)
        if self._bw_host is None:
            self._bw_setup(shape, t.dtype if t is not None else torch.from_numpy(np.zeros(1, dtype=data.dtype)).dtype)
        elif shape != self._bw_frame_shape:
            raise ValueError('writer received a frame of shape {} after frames of shape {}'.format(shape, self._bw_frame_shape))
        slot = self._bw_host[self._bw_cur][self._bw_fill]
        if t is not None and t.is_cuda:
            self._bw_enqueue_cuda(t, slot)
        elif t is not None:
            slot.copy_(t.detach().to(self._bw_dtype))
        else:
            slot.copy_(torch.from_numpy(np.ascontiguousarray(data, dtype=_np_dtype(self._bw_dtype))))
        self._bw_fill += 1
        self._buffer.append(slot)
        self._counter += 1
        self._chunk_bytes += self._estimate_frame_bytes(slot)
        logical = self._should_flush()
        if logical or self._bw_fill == self._bw_cap:
            self._bw_submit(close_chunk=logical)

    def flush_chunk(self):
        """Close the current chunk: everything received so far is handed to the writer thread."""
        if self._bw_host is None or (self._bw_fill == 0 and not self._bw_open_chunk):
            return
        self._bw_submit(close_chunk=True)

    def finalize(self):
        """Flush, wait for the background writes, aggregate the chunks into the final file."""
        self.flush_chunk()
        self._bw_drain()
        self._aggregate(self._bw_chunk_files)
        self._cleanup_chunks()

    def wait(self):
        self.finalize()

    def _should_flush(self):
        if self._every_N is not None:
            return self._counter % self._every_N == 0
        return self._chunk_bytes >= self._max_chunk_mb * self._MB

    def _estimate_frame_bytes(self, data):
        if isinstance(data, torch.Tensor):
            return data.element_size() * data.numel()
        a = np.asarray(data)
        return a.dtype.itemsize * a.size

    def _chunk_path(self, idx):
        return '{0}__chunk_{1:06d}.npy'.format(self._tmp_prefix(), idx)

    def _tmp_prefix(self):
        return self._fname

    def _write_chunk(self, chunk, idx):
        """Default (NPY) chunk writer; returns the chunk path.  Called on the writer thread, in order."""
        path = self._chunk_path(idx)
        fdir = os.path.split(path)[0]
        if fdir and not os.path.exists(fdir):
            os.makedirs(fdir)
        np.save(path, chunk)
        return path

    def _aggregate(self, chunk_files):
        """Default (NPY) aggregation through a memory-mapped output (basewriter.py:140-164)."""
        target = self._final_path()
        if len(chunk_files) == 0:
            np.save(target, np.empty((0,), dtype=np.float32))
            return
        shapes = [np.load(f, mmap_mode='r').shape for f in chunk_files]
        ntot = int(sum(s[0] for s in shapes))
        first = np.load(chunk_files[0], mmap_mode='r')
        out = np.lib.format.open_memmap(target, mode='w+', dtype=first.dtype, shape=(ntot,) + shapes[0][1:])
        pos = 0
        for f in chunk_files:
            block = np.load(f, mmap_mode='r')
            out[pos:pos + block.shape[0]] = block
            pos += block.shape[0]
        out.flush()
        del out
        print('saving file {0}'.format(target))

    def _final_path(self):
        fname = self._fname
        return fname if fname.endswith('.npy') else '{0}.npy'.format(fname)

    def _cleanup_chunks(self):
        for f in self._bw_chunk_files:
            if f is not None and os.path.exists(f):
                os.remove(f)
        self._bw_chunk_files = []

    # ---- pipeline -------------------------------------------------------------------------------
    def _bw_setup(self, shape, dtype):
        self._bw_frame_shape = shape
        self._bw_dtype = dtype
        frame_bytes = max(1, int(np.prod(shape, dtype=np.int64)) * torch.empty((), dtype=self._bw_dtype).element_size())
        if self._every_N is not None:
            want = int(self._every_N)
        else:
            want = int(-(-int(self._max_chunk_mb * self._MB) // frame_bytes))
        self._bw_cap = max(1, min(max(want, 1), max(1, _PINNED_CAP_BYTES // frame_bytes)))
        pin = torch.cuda.is_available()
        self._bw_host = [torch.empty((self._bw_cap,) + shape, dtype=self._bw_dtype, pin_memory=pin) for _ in range(2)]
        self._bw_free = threading.Semaphore(1)          # the second buffer; the first one is in use
        self._bw_queue = queue.Queue()
        self._bw_thread = threading.Thread(target=self._bw_worker, name='pdx-writer', daemon=True)
        self._bw_thread.start()

    def _bw_enqueue_cuda(self, t, slot):
        dev = t.device
        if self._bw_stage is None:
            self._bw_stage = [torch.empty(self._bw_frame_shape, dtype=self._bw_dtype, device=dev) for _ in range(2)]
            self._bw_stage_done = [torch.cuda.Event(), torch.cuda.Event()]
            self._bw_copy_stream = torch.cuda.Stream(device=dev)
            for ev in self._bw_stage_done:
                ev.record(self._bw_copy_stream)
        i = self._bw_stage_idx
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(self._bw_stage_done[i])           # the staging buffer's previous transfer has left the device
        self._bw_stage[i].copy_(t.detach().to(self._bw_dtype).reshape(self._bw_frame_shape))   # snapshot (D2D)
        self._bw_copy_stream.wait_stream(cur)
        with torch.cuda.stream(self._bw_copy_stream):
            slot.copy_(self._bw_stage[i], non_blocking=True)
            self._bw_stage_done[i].record(self._bw_copy_stream)
        self._bw_last_copy = self._bw_stage_done[i]
        self._bw_stage_idx = i ^ 1

    def _bw_submit(self, close_chunk):
        """Hand the filled part of the current pinned buffer to the writer thread and switch buffers."""
        ev = None
        if self._bw_copy_stream is not None:
            ev = torch.cuda.Event()
            ev.record(self._bw_copy_stream)
        if self._bw_fill > 0:
            # idx handed to _write_chunk numbers the pinned buffers written out; a chunk larger than the
            # pinned buffer simply becomes several consecutive pieces
            self._bw_queue.put((self._bw_cur, self._bw_fill, ev, self._bw_pieces))
            self._bw_pieces += 1
            self._bw_free.acquire()                     # wait (rarely) until the other buffer has been written out
            self._bw_cur ^= 1
            self._bw_fill = 0
        self._bw_open_chunk = not close_chunk
        if close_chunk:
            self._chunk_index += 1
            self._chunk_bytes = 0
            self._buffer = []

    def _bw_worker(self):
        while True:
            item = self._bw_queue.get()
            if item is None:
                self._bw_queue.task_done()
                return
            buf, n, ev, idx = item
            try:
                if self._bw_error is None:
                    if ev is not None:
                        ev.synchronize()
                    path = self._write_chunk(self._bw_host[buf][:n].numpy(), idx)
                    if path is not None and path not in self._bw_chunk_files:
                        self._bw_chunk_files.append(path)
            except BaseException as e:      # reported on the caller's thread
                self._bw_error = e
            finally:
                self._bw_free.release()
                self._bw_queue.task_done()

    def _bw_drain(self):
        if self._bw_queue is not None:
            self._bw_queue.join()
        self._raise_pending()

    def _raise_pending(self):
        if self._bw_error is not None:
            e, self._bw_error = self._bw_error, None
            raise e

    def _bw_shutdown(self):
        if self._bw_thread is not None and self._bw_thread.is_alive():
            self._bw_queue.put(None)
            self._bw_thread.join(timeout=5)
        self._bw_thread = None


def _np_dtype(torch_dtype):
    return torch.empty((), dtype=torch_dtype).numpy().dtype
