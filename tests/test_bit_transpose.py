"""Model of `warp_transpose32` (pyjess_b200/csrc/kernels.cuh): the five butterfly exchanges `k_classify` uses to turn
"the specs each of 32 atoms matches" into "the atoms matching each of 32 specs".  The kernel's formula, lane by lane,
must be the bit-matrix transpose: result[b] bit a == input[a] bit b."""
import random

M32 = 0xFFFFFFFF


def warp_transpose32(words):
    x = list(words)
    m = 0x0000FFFF
    j = 16
    while j > 0:
        y = [x[lane ^ j] for lane in range(32)]                     # __shfl_xor_sync(x, j)
        for lane in range(32):
            if (lane & j) == 0:
                x[lane] ^= ((((x[lane] >> j) ^ y[lane]) & m) << j) & M32
            else:
                x[lane] ^= ((y[lane] >> j) ^ x[lane]) & m
        m = (m ^ (m << (j >> 1))) & M32
        j >>= 1
    return x


def test_butterfly_is_the_bit_matrix_transpose():
    rng = random.Random(7)
    cases = [[rng.getrandbits(32) for _ in range(32)] for _ in range(50)]
    cases += [[0] * 32, [M32] * 32, [1 << i for i in range(32)], [M32 if i == 5 else 0 for i in range(32)]]
    for words in cases:
        out = warp_transpose32(words)
        for a in range(32):
            for b in range(32):
                assert (out[b] >> a) & 1 == (words[a] >> b) & 1
