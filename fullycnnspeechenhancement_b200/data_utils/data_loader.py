"""AudioParser / DataSet / DataLoader with the reference's interface
(data_utils/data_loader.py of the reference), feeding the CUDA path.

The hot-path pieces are AudioParser.parse_audio (STFT kernel) and the batch layout contract of
DataLoader.padding_batch: zero-pad every [F, T_i] spectrogram to T_max and return
[N, T_max, F, 1] (data_loader.py:198-209).  Manifest handling and noise mixing follow the
reference's formulas; file decoding uses scipy (audio_io.py) because librosa is not available."""
import codecs
import json

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
        """Mix at self.snr dB (data_loader.py:35-52): tile / crop the noise to the speech length,
        scale it so that sum(speech^2)/sum(noise^2) == 10^(snr/10)."""
        if len(speech) >= len(noise):
            reps = int(np.ceil((len(speech) - len(noise)) / len(noise)))
            for _ in range(reps):
                noise = np.concatenate((noise, noise * np.random.uniform(0, 2)))
            noise = noise[:len(speech)]
        else:
            start = np.random.randint(0, len(noise) - len(speech))
            noise = noise[start:start + len(speech)]
        p_sig = np.sum(abs(speech) ** 2)
        p_back = np.sum(abs(noise) ** 2)
        return speech + np.sqrt(p_sig / (10 ** (self.snr / 10)) / p_back) * noise

    def parse_audio(self, sig):
        return self.extractor.compute_spectrogram(sig, self.sample_rate, window_s=self.window_s,
                                                  stride_s=self.stride_s, nfft=256, use_complex=self.complex)


class DataSet(AudioParser):
    """JSON-lines manifest reader (data_loader.py:64-125): items need ``audio_filepath`` and
    ``duration``; items outside [min_duration, max_duration] are dropped."""

    def __init__(self, manifest_filepath, noise_manifest, sample_rate=16000, window_ms=32, stride_ms=16, snr=0,
                 min_duration=0.4, max_duration=float("inf"), use_complex=False):
        super(DataSet, self).__init__(sample_rate, window_ms, stride_ms, snr, None, use_complex)
        self.item_list = self._read_manifest(manifest_filepath, min_duration, max_duration)
        self.noise_list = self._read_manifest(noise_manifest, 0.0, float("inf")) if noise_manifest else None
        self.size = len(self.item_list)

    @staticmethod
    def _read_manifest(path, min_duration, max_duration):
        items = []
        with codecs.open(path, "r", "utf-8") as f:
            for line in f:
                line = line.strip()
                if not line:
                    continue
                item = json.loads(line)
                if min_duration <= float(item.get("duration", min_duration)) <= max_duration:
                    items.append(item)
        return items

    def __len__(self):
        return self.size

    def __getitem__(self, index):
        """((mix_sig, clean_sig), (mix_spec, clean_spec)) like the reference's item tuple."""
        item = self.item_list[index]
        clean, _ = self.load_audio(item.get("audio_filepath", item.get("clean_filepath")))
        if self.noise_list:
            noise_item = self.noise_list[np.random.randint(0, len(self.noise_list))]
            noise, _ = self.load_audio(noise_item["audio_filepath"])
            mix = self.add_noise(clean, noise).astype(np.float32)
        elif "noisy_filepath" in item:
            mix, _ = self.load_audio(item["noisy_filepath"])
        else:
            mix = clean
        return (mix, clean), (self.parse_audio(mix), self.parse_audio(clean))

    def shuffle(self):
        np.random.shuffle(self.item_list)


class DataLoader(object):
    def __init__(self, dataset, batch_size, sampler=None, num_works=1):
        self.dataset = dataset
        self.batch_size = batch_size
        self.sampler = sampler
        self.num_works = num_works
        n = len(dataset)
        self.bins = [list(range(i, min(i + batch_size, n))) for i in range(0, n, batch_size)]

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

    def __iter__(self):
        groups = self.sampler if self.sampler is not None else self.bins
        for index_list in groups:
            yield self.collect_fn([self.dataset[i] for i in index_list])

    def __len__(self):
        return len(self.sampler) if self.sampler is not None else len(self.bins)

    def shuffle(self):
        self.dataset.shuffle()
