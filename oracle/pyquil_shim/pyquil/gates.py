def _gate(name):
    def g(*a, **k):
        raise NotImplementedError(f"pyquil shim: gate {name} is a placeholder")
    g.__name__ = name
    return g


for _n in ["I", "X", "Y", "Z", "H", "S", "T", "RX", "RY", "RZ", "CZ", "CNOT", "XY", "MEASURE",
           "RESET", "PHASE", "CPHASE", "SWAP", "ISWAP", "CCNOT", "Gate", "QUANTUM_GATES"]:
    globals()[_n] = _gate(_n)
