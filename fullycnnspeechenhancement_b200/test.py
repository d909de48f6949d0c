"""Batched evaluation entry point, command-line compatible with the reference's test.py:

    python -m fullycnnspeechenhancement_b200.test --cfg CFG [--num-works N]

The cfg keys read here are the reference's ([data] window_ms / stride_ms / sample_rate / snr /
test_manifest_path / optional test_noise_manifest, [testing] batch_size); everything else is read by
FullyCNNTester itself.  The loop runs waveform -> waveform on the GPU (K1 -> K2 -> K3)."""
import argparse

from .config import load_conf_info
from .data_utils.data_loader import DataLoader, DataSet
from .model_utils.tester import FullyCNNTester


def _data_option(config, key, cast=str, default=None):
    """[data] option, or ``default`` when the key is absent and a default was given."""
    if default is not None and not config.has_option("data", key):
        return default
    return cast(config.get("data", key))


def build_loader(config, num_works):
    """DataSet + DataLoader exactly as the reference's main() builds them (complex spectrograms, no sampler)."""
    noise_manifest = config.get("data", "test_noise_manifest") if config.has_option("data", "test_noise_manifest") else None
    dataset = DataSet(
        manifest_filepath=_data_option(config, "test_manifest_path"),
        noise_manifest=noise_manifest,
        sample_rate=_data_option(config, "sample_rate", int),
        window_ms=_data_option(config, "window_ms", int),
        stride_ms=_data_option(config, "stride_ms", int),
        snr=_data_option(config, "snr", float),
        use_complex=True,
    )
    return DataLoader(dataset, int(config.get("testing", "batch_size")), sampler=None, num_works=num_works)


def main(config, num_works):
    loader = build_loader(config, num_works)
    tester = FullyCNNTester(config)
    return tester.test(loader)


if __name__ == "__main__":
    cli = argparse.ArgumentParser(description="Testing")
    cli.add_argument("--cfg", default="", type=str, help="cfg file for test")
    cli.add_argument("--num-works", default=16, type=int, help="threads fetching the items of a batch")
    opts = cli.parse_args()
    main(load_conf_info(opts.cfg), opts.num_works)
