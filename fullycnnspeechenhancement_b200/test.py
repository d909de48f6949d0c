"""Batched evaluation entry point with the reference's interface (test.py of the reference):
python -m fullycnnspeechenhancement_b200.test --cfg CFG --num-works N"""
import argparse

from .config import load_conf_info
from .data_utils.data_loader import DataLoader, DataSet
from .model_utils.tester import FullyCNNTester


def main(config, num_works):
    window_ms = int(config.get("data", "window_ms"))
    stride_ms = int(config.get("data", "stride_ms"))
    sample_rate = int(config.get("data", "sample_rate"))
    noise = config.get("data", "test_noise_manifest") if config.has_option("data", "test_noise_manifest") else None
    dataset = DataSet(manifest_filepath=config.get("data", "test_manifest_path"), noise_manifest=noise,
                      sample_rate=sample_rate, window_ms=window_ms, stride_ms=stride_ms,
                      snr=float(config.get("data", "snr")), use_complex=True)
    loader = DataLoader(dataset, int(config.get("testing", "batch_size")), sampler=None, num_works=num_works)
    return FullyCNNTester(config).test(loader)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Testing")
    ap.add_argument("--cfg", default="", type=str, help="cfg file for test")
    ap.add_argument("--num-works", default=16, type=int, help="kept for command-line compatibility")
    a = ap.parse_args()
    main(load_conf_info(a.cfg), a.num_works)
