"""Potential-energy model (`potential.GravMagModule`) and HMC sampler (`hmc.HMCSample`)."""
