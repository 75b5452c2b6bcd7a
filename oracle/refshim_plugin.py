"""ORACLE - TEST INFRASTRUCTURE ONLY.  pytest plugin: `-p oracle.refshim_plugin` installs the fake
mujoco/gymnasium modules so the reference's own test-suite can run on the CPU oracle."""
from oracle import fake_mujoco

fake_mujoco.install()
