"""Host-side re-statement of the library's stateless dropout hash (csrc/stages.cuh::dropout_hash), used by
tests to check the kept fraction without reading device activations."""


def dropout_hash(seed: int, idx: int) -> int:
    m = (1 << 64) - 1
    z = (seed + idx * 0x9E3779B97F4A7C15) & m
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    z ^= z >> 31
    return z >> 32


def kept_fraction(p: float, seed: int, n: int) -> float:
    thr = int(p * 4294967296.0)
    return sum(dropout_hash(seed, i) >= thr for i in range(n)) / n
