"""qmcb200: a B200-native (sm_100a) walker-ensemble QMC engine for the
multi-rods Bijl-Jastrow Bose gas (the ``mrbp_qmc`` model of PhD-QMCLib).

``model``   host-side model spec (parameter derivation)
``engine``  thin object wrapper over the C ABI of ``libqmcb200.so``
``dmc``     DMC sampler with the reference's ``Sampling`` interface
``vmc``     VMC sampler with the reference's ``Sampling`` interface
"""
from . import model  # noqa: F401

__version__ = '0.1.0'
