"""Packaging: ``pip install -e .`` builds the in-tree native libraries and installs the ``epl-launch`` console script
(reference ``setup.py:51-97``: wheel with ``libcommunicators.so`` as package data + ``epl-launch``)."""
from setuptools import find_packages, setup
from setuptools.command.build_py import build_py


class BuildNative(build_py):
  def run(self):
    from easyparallellibrary_b200.build import build_all
    build_all()
    super().run()


setup(
    name="easyparallellibrary-b200",
    version="0.1.0",
    packages=find_packages(include=["easyparallellibrary_b200*", "epl*"]),
    package_data={"easyparallellibrary_b200": ["lib/*.so", "csrc/*"]},
    cmdclass={"build_py": BuildNative},
    entry_points={"console_scripts": ["epl-launch = easyparallellibrary_b200.utils.launcher:main"]},
    python_requires=">=3.10",
)
