"""Restated from pyquil's published conventions (pyquil.simulation.matrices)."""
import numpy as np

I = np.array([[1.0, 0.0], [0.0, 1.0]])
X = np.array([[0.0, 1.0], [1.0, 0.0]])
Y = np.array([[0.0, 0.0 - 1.0j], [0.0 + 1.0j, 0.0]])
Z = np.array([[1.0, 0.0], [0.0, -1.0]])
H = (1.0 / np.sqrt(2.0)) * np.array([[1.0, 1.0], [1.0, -1.0]])
CNOT = np.array([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
QUANTUM_GATES = {"I": I, "X": X, "Y": Y, "Z": Z, "H": H, "CNOT": CNOT}

_s2, _s3 = np.sqrt(2.0), np.sqrt(3.0)
STATES = {
    "X": [np.array([1, 1]) / _s2, np.array([1, -1]) / _s2],
    "Y": [np.array([1, 1j]) / _s2, np.array([1, -1j]) / _s2],
    "Z": [np.array([1, 0]), np.array([0, 1])],
    "SIC": [
        np.array([1, 0]),
        np.array([1, _s2]) / _s3,
        np.array([1, np.exp(-2j * np.pi / 3) * _s2]) / _s3,
        np.array([1, np.exp(2j * np.pi / 3) * _s2]) / _s3,
    ],
}
