"""AudioParser / DataSet / Sampler / DataLoader with the reference's interface
(data_utils/data_loader.py of the reference), feeding the CUDA path.

The hot-path pieces are AudioParser.parse_audio (STFT kernel) and the batch layout contract of
DataLoader.padding_batch: zero-pad every [F, T_i] spectrogram to T_max and return
[N, T_max, F, 1] (data_loader.py:198-209).  Manifest handling, clean/noise pairing, noise mixing
and batch binning follow the reference's behaviour; file decoding uses scipy (audio_io.py)
because librosa is not available, and items are fetched by a thread pool (``num_works``) where
the reference forks joblib workers (the per-item work here is file decoding plus a CUDA call)."""
import codecs
import json
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .. import audio_io
from .audio_feature import AudioFeature


class AudioParser(object):
    def __init__(self, sample_rate=8000, window_ms=32, stride_ms=16, snr=0, windows_name=None, use_complex=False):
        self.snr = snr
        self.sample_rate = sample_rate
        self.window_s = window_ms / 1000
        self.stride_s = stride_ms / 1000
        self.extractor = AudioFeature(windows_name)
        self.complex = use_complex

    def load_audio(self, audio_filepath):
        return audio_io.load_wav(audio_filepath, self.sample_rate)

    def add_noise(self, speech, noise):
        """Mix at self.snr dB (data_loader.py:35-52).  A noise clip shorter than the speech is
        extended by doubling (each appended copy of the whole clip-so-far scaled by one
        uniform(0, 2) draw), a longer one is cropped at a random start; then it is scaled so that
        sum(speech^2) / sum(noise^2) == 10^(snr/10).  Draws come from numpy's global generator in
        the reference's order, so a seeded run mixes identically."""
        n_speech, n_noise = len(speech), len(noise)
        if n_speech < n_noise:
            first = np.random.randint(0, n_noise - n_speech)
            bed = noise[first:first + n_speech]
        else:
            bed = noise
            doublings = int(np.ceil((n_speech - n_noise) / n_noise))   # of the ORIGINAL length, as in the reference
            for _ in range(doublings):
                gain = np.random.uniform(0, 2)
                bed = np.concatenate((bed, bed * gain))
            bed = bed[:n_speech]
        speech_energy = np.sum(abs(speech) ** 2)
        noise_energy = np.sum(abs(bed) ** 2)
        target_noise_energy = speech_energy / (10 ** (self.snr / 10))
        return speech + np.sqrt(target_noise_energy / noise_energy) * bed

    def parse_audio(self, sig):
        return self.extractor.compute_spectrogram(sig, self.sample_rate, window_s=self.window_s,
                                                  stride_s=self.stride_s, nfft=256, use_complex=self.complex)


class DataSet(AudioParser):
    """JSON-lines manifests (data_loader.py:64-125).  Every line needs ``duration``; lines outside
    [min_duration, max_duration] are dropped.  With a noise manifest, item i is
    ``audio_filepath`` of the speech manifest mixed with ``audio_filepath`` of noise item i (the
    noise list is repeated until it is at least as long); without one, every line names a
    ``clean_audio_filepath`` / ``mix_audio_filepath`` pair."""

    def __init__(self, manifest_filepath, noise_manifest, sample_rate=16000, window_ms=32, stride_ms=16, snr=0,
                 min_duration=0.4, max_duration=float("inf"), windows_name=None, use_complex=False):
        super(DataSet, self).__init__(sample_rate=sample_rate, window_ms=window_ms, stride_ms=stride_ms, snr=snr,
                                      windows_name=windows_name, use_complex=use_complex)
        self.min_duration = min_duration
        self.max_duration = max_duration
        self.noise_manifest = noise_manifest
        self.item_list = self.read_manifest(manifest_filepath)
        if noise_manifest is not None:
            self.noise_list = self.read_manifest(noise_manifest)
            if len(self.noise_list) < len(self.item_list):
                self.noise_list = self.noise_list * int(np.ceil(len(self.item_list) / len(self.noise_list)))
            assert len(self.noise_list) >= len(self.item_list)

    def read_manifest(self, manifest_path):
        manifest = []
        with codecs.open(manifest_path, "r", "utf-8") as f:
            for line in f:
                try:
                    item = json.loads(line)
                except Exception as e:
                    raise IOError("Error reading manifest: %s" % str(e))
                if self.max_duration >= item["duration"] >= self.min_duration:
                    manifest.append(item)
        return manifest

    def item_name(self, index):
        """File the outputs of item ``index`` are named after.  The reference's tester always reads
        ``audio_filepath`` (tester.py:148), which paired manifests do not have; here those fall
        back to the clean file."""
        item = self.item_list[index]
        return item["audio_filepath"] if "audio_filepath" in item else item["clean_audio_filepath"]

    def load_pair(self, index):
        """(mix_sig, speech) of item ``index`` -- the waveform half of ``__getitem__``."""
        item = self.item_list[index]
        if self.noise_manifest is not None:
            speech, _ = self.load_audio(item["audio_filepath"])
            noise, _ = self.load_audio(self.noise_list[index]["audio_filepath"])
            return self.add_noise(speech, noise), speech
        speech, _ = self.load_audio(item["clean_audio_filepath"])
        mix_sig, _ = self.load_audio(item["mix_audio_filepath"])
        return mix_sig, speech

    def __getitem__(self, index):
        """((mix_sig, speech), (mix_spec, speech_spec)), the reference's item tuple."""
        mix_sig, speech = self.load_pair(index)
        speech_spec = self.parse_audio(speech)
        mix_spec = self.parse_audio(mix_sig)
        return (mix_sig, speech), (mix_spec, speech_spec)

    def __len__(self):
        return len(self.item_list)

    def __call__(self, *args, **kwargs):
        return self

    def shuffle(self):
        np.random.shuffle(self.item_list)


class Sampler(object):
    """Batch index bins in random order (data_loader.py:128-160).  ``drop_last`` cuts the ragged
    tail off the dataset's item list; otherwise the list is EXTENDED with its last items up to the
    next multiple of the batch size -- a whole extra batch when it already divides, as in the
    reference (int(n / b) + 1 batches).  The dataset's ``item_list`` is modified in place, which is
    what makes ``len(dataset)`` a multiple of the batch size afterwards."""

    def __init__(self, dataset, batch_size, start_index=0, drop_last=False):
        self.dataset, self.batch_size, self.start_index = dataset, batch_size, start_index
        items = self.dataset.item_list
        count = len(items)
        if drop_last:
            tail = count % batch_size
            # a zero tail gives items[:-0] == []: the reference empties the list in that case, so do we
            self.dataset.item_list = items[:-tail]
        else:
            missing = (count // batch_size + 1) * batch_size - count
            items.extend(items[-missing:])
        total = len(self.dataset)
        self.bins = [list(range(lo, min(lo + batch_size, total))) for lo in range(0, total, batch_size)]
        order = np.random.permutation(len(self.bins) - start_index) + start_index
        self.indices = order.tolist()

    def __iter__(self):
        for which in self.indices:
            members = self.bins[which]
            np.random.shuffle(members)       # in place, like the reference: the bin stays shuffled
            yield members

    def __len__(self):
        return len(self.bins) - self.start_index

    def iter_num(self):
        return len(self.indices)

    def reset_start_index(self, start_index):
        self.start_index = start_index

    def __call__(self, *args, **kwargs):
        return self


class DataLoader(object):
    def __init__(self, dataset, batch_size, sampler=None, num_works=2):
        self.dataset = dataset
        self.batch_size = batch_size
        self.sampler = sampler
        self.num_works = num_works
        self.q = []
        if self.sampler is None:
            n = len(self.dataset)
            self.bins = [list(range(i, min(i + self.batch_size, n))) for i in range(0, n, self.batch_size)]

    def one_point(self, x):
        a, b = self.dataset[x]
        return a, b

    def pool_process(self, index_list):
        """Fetch the items of one batch in order.  Items that draw random numbers (noise mixing)
        are fetched serially so that a seeded run reproduces; pure file pairs go through a thread
        pool of ``num_works``."""
        serial = getattr(self.dataset, "noise_manifest", None) is not None or self.num_works is None or self.num_works <= 1
        if serial:
            results = [self.one_point(i) for i in index_list]
        else:
            with ThreadPoolExecutor(max_workers=int(self.num_works)) as pool:
                results = list(pool.map(self.one_point, index_list))
        self.q.extend(results)

    @staticmethod
    def padding_batch(batch_list):
        """data_loader.py:198-209: [F,T_i] list -> zero padded [N, T_max, F, 1]."""
        t_max = max(arr.shape[1] for arr in batch_list)
        out = np.zeros((len(batch_list), t_max, batch_list[0].shape[0], 1), dtype=batch_list[0].dtype)
        for i, arr in enumerate(batch_list):
            out[i, :arr.shape[1], :, 0] = np.transpose(arr)
        return out

    def collect_fn(self, batch_data):
        mix_sig = [b[0][0] for b in batch_data]
        clean_sig = [b[0][1] for b in batch_data]
        mix = self.padding_batch([b[1][0] for b in batch_data])
        clean = self.padding_batch([b[1][1] for b in batch_data])
        assert mix.shape == clean.shape
        return mix, clean, mix_sig, clean_sig

    def _index_groups(self):
        if self.sampler is None:
            return self.bins
        if self.sampler.batch_size != self.batch_size:
            self.batch_size = self.sampler.batch_size
            print("Warrning: sampler.batch_size != batch_size. batch_size changed!")
        return self.sampler

    def __iter__(self):
        for group in self._index_groups():
            self.pool_process(group)
            fetched, self.q = self.q, []
            yield self.collect_fn(fetched)

    def __len__(self):
        return len(self.sampler) if self.sampler is not None else len(self.bins)

    def shuffle(self):
        self.dataset.shuffle()
