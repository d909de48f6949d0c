"""Entry-point shim with the reference's file name: see fullycnnspeechenhancement_b200/test.py."""
from fullycnnspeechenhancement_b200.test import *  # noqa: F401,F403

if __name__ == "__main__":
    import runpy
    runpy.run_module("fullycnnspeechenhancement_b200.test", run_name="__main__")
