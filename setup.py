"""Packaging: `pip install -e .` builds the sm_100a native libraries in-tree first (native/build.py drives nvcc/g++).

Reference counterpart: install.sh + Dockerfiles (/root/reference/install.sh:1-359) only set up a Python venv;
here the native runtime (collectives, GEMM, fused ops, stager, task runner, probe, mpibench, diskbench) is part of the package.
"""
import importlib.util
import os

from setuptools import setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildNativeThenPy(build_py):
    def run(self):
        if os.environ.get("SHIPYARD_SKIP_NATIVE") != "1":
            spec = importlib.util.spec_from_file_location("shipyard_native_build", os.path.join(ROOT, "native", "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod.build_all(None, force=False, quiet=False)
        super().run()


setup(cmdclass={"build_py": BuildNativeThenPy})
