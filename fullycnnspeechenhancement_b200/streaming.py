"""Online (block-by-block) enhancement on top of the batched GPU path -- the "real-time" item of the
reference's to-do list (readme.md:76-79), built on the chunking rules of ``Enhancer.enhance_stream``.

The reference has no streaming mode: it enhances a whole file at once.  Its computation is local,
though: frames interact only through the 8-tap time kernel of the first layer
(3 frames back, 4 ahead) and through two one-pole recurrences (pre-emphasis, de-emphasis).  A block
of the signal can therefore be enhanced exactly as in the whole-file run if it is processed together
with

* a look-back of 13 hops (1,664 samples): 5 hops until the network output is exact, 8 more for the
  de-emphasis carry to decay by 0.97^1024 ~ 3e-14, and
* a look-ahead of 6 hops (768 samples): the output segment of the last kept sample needs 4 further
  complete frames.

``StreamingEnhancer.push`` accepts blocks of any size and returns the enhanced samples that have become
final; ``flush`` returns the tail at the end of the stream.  The algorithmic latency is the look-ahead
plus the block granularity: 768 samples + ``block`` samples at 8 kHz (0.16 s with 512-sample blocks).
The output equals the whole-file result to float32 rounding (> 100 dB; tests/test_gpu_parity.py)."""
import numpy as np

FRAME_HOP = 128
LOOK_BACK = 13 * FRAME_HOP
LOOK_AHEAD = 6 * FRAME_HOP


class StreamingEnhancer(object):
    def __init__(self, enhancer, block=4096):
        """``enhancer``: anything with ``enhance(list_of_waveforms) -> list_of_waveforms`` (engine.Enhancer).
        ``block``: samples enhanced per GPU call, rounded down to a multiple of the 128-sample hop."""
        if block < FRAME_HOP:
            raise ValueError("block must be at least one hop (128 samples)")
        self.enhancer = enhancer
        self.block = block // FRAME_HOP * FRAME_HOP
        self._buf = np.zeros(0, np.float32)   # samples from `_buf0` on (absolute index of _buf[0])
        self._buf0 = 0
        self._done = 0                        # absolute index of the first sample not yet returned
        self._total = 0                       # samples pushed so far

    @property
    def latency_samples(self):
        return LOOK_AHEAD + self.block

    def _emit(self, end, final):
        """Enhance [self._done, end) with its halo and drop what is no longer needed."""
        outs = []
        while self._done < end:
            e = min(end, self._done + self.block) if not final else min(end, self._done + max(self.block, 16 * FRAME_HOP))
            a = max(0, self._done - LOOK_BACK)
            b = min(self._total, e + LOOK_AHEAD)
            piece = self._buf[a - self._buf0:b - self._buf0]
            res = self.enhancer.enhance([piece])[0]
            outs.append(np.asarray(res[self._done - a:e - a], dtype=np.float32))
            self._done = e
        keep_from = max(0, self._done - LOOK_BACK)
        if keep_from > self._buf0:
            self._buf = self._buf[keep_from - self._buf0:]
            self._buf0 = keep_from
        return np.concatenate(outs) if outs else np.zeros(0, np.float32)

    def push(self, samples):
        """Append ``samples`` (1-D float) to the stream; returns the enhanced samples that are final now
        (possibly empty).  Blocks always start on the hop grid of the whole signal."""
        x = np.asarray(samples, dtype=np.float32).reshape(-1)
        self._buf = np.concatenate([self._buf, x])
        self._total += len(x)
        # a block [done, done + block) is final once its look-ahead has arrived
        n_blocks = (self._total - LOOK_AHEAD - self._done) // self.block
        if n_blocks <= 0:
            return np.zeros(0, np.float32)
        return self._emit(self._done + n_blocks * self.block, final=False)

    def flush(self):
        """End of the stream: returns everything not yet returned (the signal's own end replaces the look-ahead)."""
        out = self._emit(self._total, final=True)
        self._buf = np.zeros(0, np.float32)
        self._buf0 = self._done = self._total = 0
        return out
