"""Host-side re-statement of the library's stateless dropout hash (csrc/stages.cuh::dropout_bits), used by
tests to check the kept fraction without reading device activations."""


def dropout_bits(seed: int, group: int) -> int:
    m = 0xFFFFFFFF
    x = (group & m) ^ (seed & m)
    y = ((group >> 32) & m) ^ ((seed >> 32) & m)
    x ^= (y * 0x9E3779B9) & m
    x ^= x >> 16
    x = (x * 0x7feb352d) & m
    x ^= x >> 15
    x = (x * 0x846ca68b) & m
    x ^= x >> 16
    return x


def kept_fraction(p: float, seed: int, n: int) -> float:
    thr = int(p * 256.0 + 0.5)
    kept = 0
    for i in range(n):
        kept += ((dropout_bits(seed, i >> 2) >> (8 * (i & 3))) & 255) >= thr
    return kept / n
