"""float64 numpy restatement of the reference STFT feature path (test oracle).

Follows /root/reference/data_utils/audio_feature.py line by line; every function
cites the lines it restates.  Written without ``np.mat`` (removed in numpy 2).
"""
import numpy as np

PRE_EMPHASIS = 0.97  # audio_feature.py:53


def pre_emphasis(signal):
    """audio_feature.py:46-55.  ``signal`` keeps its dtype: for the float32 arrays
    librosa hands the reference, ``0.97`` is a weak python scalar, so the product
    and the difference are both rounded to float32 (no fused multiply-add)."""
    signal = np.asarray(signal)
    return np.append(signal[0], signal[1:] - PRE_EMPHASIS * signal[:-1])


def frame_count(signal_length, frame_length=256, frame_step=128):
    """audio_feature.py:67-70: ``ceil(|L - fl| / fs + 1)`` (note the abs())."""
    return int(np.ceil(float(np.abs(signal_length - frame_length)) / frame_step + 1))


def en_frame(frame_size, frame_stride, sample_rate, signal):
    """audio_feature.py:57-77: zero-pad to T*fs+fl samples (float64 zeros, so the
    result is float64) and gather frame t = samples [t*fs, t*fs+fl)."""
    frame_length = int(round(frame_size * sample_rate))
    frame_step = int(round(frame_stride * sample_rate))
    signal_length = len(signal)
    num_frames = frame_count(signal_length, frame_length, frame_step)
    pad_signal_length = num_frames * frame_step + frame_length
    pad_signal = np.append(signal, np.zeros(pad_signal_length - signal_length))
    starts = np.arange(num_frames, dtype=np.int64) * frame_step
    indices = starts[:, None] + np.arange(frame_length, dtype=np.int64)[None, :]
    return frame_length, pad_signal[indices]


def frame_indices(signal_length, frame_length=256, frame_step=128):
    """The integer gather table of audio_feature.py:74-76 (bit-exact parity item)."""
    num_frames = frame_count(signal_length, frame_length, frame_step)
    starts = np.arange(num_frames, dtype=np.int64) * frame_step
    return starts[:, None] + np.arange(frame_length, dtype=np.int64)[None, :]


def add_windows(frame_length, frames, window=np.hamming):
    """audio_feature.py:79-88 (window is always Hamming on the shipped path:
    audio_feature.py:13-20 with windows_name=None)."""
    return frames * window(frame_length)


def compute_spectrogram(signal, sample_rate, window_s=0.032, stride_s=0.016, nfft=256,
                        use_complex=True):
    """audio_feature.py:22-44.  Returns [F, T] (complex128, or float32 magnitude)."""
    if stride_s > window_s:
        raise ValueError("Stride size must not be greater than window size.")
    emphasized = pre_emphasis(signal)
    frame_length, frames = en_frame(window_s, stride_s, sample_rate, emphasized)
    frames = add_windows(frame_length, frames)
    fft_frames = np.fft.rfft(frames, nfft)          # audio_feature.py:90-99
    if use_complex:
        return np.transpose(fft_frames)
    return np.transpose(power_spectrum(fft_frames)).astype(np.float32)


def power_spectrum(frames):
    """audio_feature.py:101-110: linear magnitude (no log, no 1/N)."""
    return np.absolute(frames)


def divide_phase(fft_frames):
    """audio_feature.py:112-115: exp(j*angle(X)); X == 0 gives 1+0j."""
    return np.exp(1.j * np.angle(fft_frames))


def padding_batch(batch_list):
    """data_loader.py:198-209: zero-pad [F,T_i] to T_max, stack, -> [N,T,F,1]."""
    max_array = max(batch_list, key=lambda x: x.shape[1])
    batch = []
    for arr in batch_list:
        sample = np.zeros_like(max_array)
        sample[:arr.shape[0], :arr.shape[1]] = arr
        batch.append(sample)
    batch = np.expand_dims(np.array(batch), axis=1)
    return np.transpose(batch, (0, 3, 2, 1))
