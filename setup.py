"""Packaging shim: the wheel carries the CUDA sources (sdaopt_b200/csrc) so that
`python -m sdaopt_b200.build` and `functors.compile_user_functor` work from an installed package;
the C-ABI header lives in include/ of the checkout and is copied next to the sources here."""
import os
import shutil

from setuptools import setup
from setuptools.command.build_py import build_py


class BuildPy(build_py):
    def run(self):
        super().run()
        here = os.path.dirname(os.path.abspath(__file__))
        dst = os.path.join(self.build_lib, "sdaopt_b200", "csrc")
        os.makedirs(dst, exist_ok=True)
        shutil.copy2(os.path.join(here, "include", "sdaopt_b200.h"), os.path.join(dst, "sdaopt_b200.h"))


setup(cmdclass={"build_py": BuildPy})
