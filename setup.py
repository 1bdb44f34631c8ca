"""`pip wheel .` / `python setup.py bdist_wheel`: builds the native parts in-tree first (build_native.py), then packages them."""
import os
import sys

from setuptools import setup
from setuptools.command.build_py import build_py
from setuptools.dist import Distribution

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildNative(build_py):
    def run(self):
        sys.path.insert(0, ROOT)
        import build_native
        build_native.build()
        super().run()


class BinaryDistribution(Distribution):
    def has_ext_modules(self):     # platform wheel: it carries an sm_100a cubin and a CPython extension
        return True


setup(cmdclass={"build_py": BuildNative}, distclass=BinaryDistribution)
