"""Entry-point shim with the reference's file name: see fullycnnspeechenhancement_b200/infer.py."""
from fullycnnspeechenhancement_b200.infer import *  # noqa: F401,F403

if __name__ == "__main__":
    import runpy
    runpy.run_module("fullycnnspeechenhancement_b200.infer", run_name="__main__")
