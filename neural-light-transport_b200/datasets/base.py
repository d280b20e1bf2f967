"""Host-side input pipeline -- the caller side of the hot path.

Mirrors the protocol of the reference's `datasets.base.Dataset` (nlt/datasets/base.py:26-117): a dataset is
constructed from `(config, mode, shuffle_buffer_size, prefetch_buffer_size, n_map_parallel_calls)`, globs its
example ids, and `build_pipeline(filter_predicate, seed, no_batch)` returns an iterable of (batched) examples:
files sorted -> optional filter -> parallel load (order preserving) -> optional cache -> post-cache hook ->
shuffle buffer (train only) -> batch of `bs` (last batch may be short) -> prefetch.

The reference builds this with tf.data; here it is a thread pool + a bounded queue, and a batch is the 11-tuple
the model consumes: python lists for the two id fields, float32 torch tensors (pinned when CUDA is present, so
the train step's host->device copies are asynchronous) for the rest.
"""
import queue
import random
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_AUTOTUNE_WORKERS = 16
_AUTOTUNE_PREFETCH = 2


class Dataset:
    def __init__(self, config, mode, shuffle_buffer_size=64, prefetch_buffer_size=None, n_map_parallel_calls=None):
        self._validate_mode(mode)
        self.config = config
        self.mode = mode
        self.shuffle_buffer_size = shuffle_buffer_size
        # None == "autotune" in the reference (tf.data.experimental.AUTOTUNE, base.py:33-38)
        self.prefetch_buffer_size = _AUTOTUNE_PREFETCH if prefetch_buffer_size is None else prefetch_buffer_size
        self.n_map_parallel_calls = _AUTOTUNE_WORKERS if n_map_parallel_calls is None else n_map_parallel_calls
        self.files = self._glob()
        assert self.files, "No files to process into a dataset"
        self.bs = self._get_batch_size()

    @staticmethod
    def _validate_mode(mode):
        allowed = ('train', 'vali', 'test')
        if mode not in allowed:
            raise ValueError("Invalid mode: {provided}. Allowed modes: {allowed}".format(provided=mode, allowed=allowed))

    def _glob(self):
        """Returns the list of example ids / paths of this mode."""
        raise NotImplementedError

    def _get_batch_size(self):
        if 'bs' not in self.config['DEFAULT'].keys():
            raise ValueError("Specify batch size either as 'bs' in the configuration file, "
                             "or override this function to generate a value another way")
        return self.config.getint('DEFAULT', 'bs')

    def _process_example_precache(self, path):
        """Loads one example; the result is what gets cached when `cache = True`."""
        raise NotImplementedError

    def _process_example_postcache(self, *args):  # pylint: disable=no-self-use
        """Per-epoch processing that must not be cached (randomness); identity by default."""
        return args

    def build_pipeline(self, filter_predicate=None, seed=None, no_batch=False, pin_memory=None):
        files = sorted(self.files)
        if filter_predicate is not None:
            files = [f for f in files if filter_predicate(f)]
        if pin_memory is None:
            pin_memory = torch.cuda.is_available()
        return DataPipe(self, files, seed=seed, no_batch=no_batch, pin_memory=pin_memory)


class DataPipe:
    """One pass over the dataset per `iter()` (like iterating a tf.data dataset): every pass reloads (or reads the
    cache), reshuffles with `seed + pass index` when training, and re-batches."""

    def __init__(self, dataset, files, seed=None, no_batch=False, pin_memory=False, limit=-1, shard=None):
        self.dataset = dataset
        self.files = files
        self.seed = seed
        self.no_batch = no_batch
        self.pin_memory = pin_memory
        self.limit = limit          # elements (batches, or examples when no_batch) per pass; -1 = all
        self.shard_spec = shard     # (world, rank): this rank's slice of every global batch
        self._cache = {}
        self._cache_lock = threading.Lock()
        self._passes = 0

    # -- tf.data-like combinators used by the drivers (trainvali.py:107-113, 90) --
    def take(self, n):
        p = DataPipe(self.dataset, self.files, self.seed, self.no_batch, self.pin_memory, limit=n, shard=self.shard_spec)
        p._cache = self._cache
        return p

    def shard(self, world, rank):
        """The reference hands the batched dataset to `strategy.experimental_distribute_dataset`, which splits
        each global batch along axis 0 across the replicas; one process per GPU takes its own slice here."""
        if not 0 <= rank < world:
            raise ValueError('rank %d outside world of %d' % (rank, world))
        p = DataPipe(self.dataset, self.files, self.seed, self.no_batch, self.pin_memory, limit=self.limit,
                     shard=(world, rank))
        p._cache = self._cache
        return p

    def __len__(self):
        n = len(self.files)
        if not self.no_batch:
            n = -(-n // self.dataset.bs)
        return n if self.limit < 0 else min(n, self.limit)

    # -- stages --
    def _load(self, f):
        ds = self.dataset
        use_cache = ds.config.getboolean('DEFAULT', 'cache', fallback=False)
        if use_cache:
            with self._cache_lock:
                hit = self._cache.get(f)
            if hit is not None:
                return ds._process_example_postcache(*hit)
        ex = ds._process_example_precache(f)
        if use_cache:
            with self._cache_lock:
                self._cache[f] = ex
        return ds._process_example_postcache(*ex)

    def _examples(self, pool):
        """Order-preserving parallel map with a bounded window of loads in flight."""
        window = max(1, self.dataset.n_map_parallel_calls)
        pending = []
        it = iter(self.files)
        for f in it:
            pending.append(pool.submit(self._load, f))
            if len(pending) >= window:
                yield pending.pop(0).result()
        for fut in pending:
            yield fut.result()

    def _shuffled(self, examples, rng):
        """tf.data shuffle semantics: keep a buffer of `shuffle_buffer_size`, emit a random slot, refill it."""
        size = max(1, int(self.dataset.shuffle_buffer_size))
        buf = []
        for ex in examples:
            buf.append(ex)
            if len(buf) > size:
                i = rng.randrange(len(buf))
                buf[i], buf[-1] = buf[-1], buf[i]
                yield buf.pop()
        while buf:
            i = rng.randrange(len(buf))
            buf[i], buf[-1] = buf[-1], buf[i]
            yield buf.pop()

    def _to_tensor(self, arr):
        arr = np.asarray(arr)
        # uint8 stays uint8 (datasets with uint8_inputs: image samples travel as bytes, the model divides on the GPU)
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.uint8 if arr.dtype == np.uint8 else np.float32))
        return t.pin_memory() if self.pin_memory else t

    def _collate(self, exs):
        cols = list(zip(*exs))
        out = []
        for col in cols:
            if isinstance(col[0], (bytes, str)):
                out.append([c if isinstance(c, bytes) else c.encode() for c in col])
            else:
                out.append(self._to_tensor(np.stack(col, axis=0)))
        return tuple(out)

    def _slice(self, batch):
        """This rank's equal share of a global batch, or None when the batch does not split evenly (a short last
        batch): every rank drops it, so the ranks keep stepping -- and all-reducing -- in lockstep."""
        if self.shard_spec is None:
            return batch
        world, rank = self.shard_spec
        n = len(batch[0])
        if n % world:
            return None
        per = n // world
        return tuple(c[rank * per:(rank + 1) * per] for c in batch)

    def _elements(self):
        is_train = self.dataset.mode == 'train'
        pass_index = self._passes
        self._passes += 1
        with ThreadPoolExecutor(max_workers=max(1, self.dataset.n_map_parallel_calls)) as pool:
            exs = self._examples(pool)
            if is_train:
                seed = None if self.seed is None else self.seed + pass_index
                exs = self._shuffled(exs, random.Random(seed))
            if self.no_batch:
                for ex in exs:
                    yield tuple(v if isinstance(v, (bytes, str)) else self._to_tensor(v) for v in ex)
                return
            cur = []
            for ex in exs:
                cur.append(ex)
                if len(cur) == self.dataset.bs:
                    el = self._slice(self._collate(cur))
                    if el is not None:
                        yield el
                    cur = []
            if cur:
                el = self._slice(self._collate(cur))
                if el is not None:
                    yield el

    def __iter__(self):
        """Prefetch: a producer thread keeps `prefetch_buffer_size` elements ready."""
        depth = max(1, int(self.dataset.prefetch_buffer_size))
        q = queue.Queue(maxsize=depth)
        stop = threading.Event()
        done = object()

        def produce():
            try:
                n = 0
                for el in self._elements():
                    if stop.is_set() or (0 <= self.limit <= n):
                        break
                    q.put(el)
                    n += 1
                q.put(done)
            except BaseException as e:   # surfaced in the consumer
                q.put(e)

        th = threading.Thread(target=produce, daemon=True)
        th.start()
        try:
            while True:
                el = q.get()
                if el is done:
                    return
                if isinstance(el, BaseException):
                    raise el
                yield el
        finally:
            stop.set()
            while th.is_alive():          # unblock a producer waiting on a full queue
                try:
                    q.get_nowait()
                except queue.Empty:
                    th.join(timeout=0.01)
