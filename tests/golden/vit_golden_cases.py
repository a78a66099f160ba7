"""Golden-vector case list shared by make_golden.py (runs the reference decoder on a B200) and the
tests that replay the vectors against the oracle (CPU) and the CUDA path (GPU).

Each case: (name, options, n_bits, seed, sigma, zero).  Inputs come from
oracle.make_channel_det (integer-only, bit-reproducible); the fixture stores the SHA-256 of the packed
input so a drifting generator is caught, and the reference decoder's output words.
"""
H, S4, S8, S16, F = 0, 1, 2, 3, 4
B32, B16, F16 = 0x00, 0x10, 0x20
O32, O16 = 0x000, 0x100

N_FULL = 6400 * 32 + 64 + 32 * 1234   # every one of the 6400 segments non-empty, ragged (P % 6400 != 0)
N_O16 = 6400 * 16 * 3 + 64 + 16 * 777 # 16-bit packs, odd pack counts -> exercises the O_B16 over-run
N_SMALL = 64 + 32 * 1500 + 5          # fewer packs than segments

CASES = [
    # name,              options,          n_bits,  seed, sigma, zero
    ("h_b32_o32",        H | B32 | O32,    N_FULL,  11,   0.70,  False),
    ("h_b32_o32_clean",  H | B32 | O32,    N_SMALL, 12,   0.00,  False),
    ("h_b32_o16",        H | B32 | O16,    N_O16,   13,   0.70,  False),
    ("h_b16_o32",        H | B16 | O32,    N_FULL,  14,   0.80,  False),
    ("h_f16_o16",        H | F16 | O16,    N_O16,   15,   0.80,  False),
    ("s4_b16_o32",       S4 | B16 | O32,   N_FULL,  16,   0.90,  False),
    ("s4_b16_o16",       S4 | B16 | O16,   N_O16,   17,   0.90,  False),
    ("s4_b32_o32",       S4 | B32 | O32,   N_FULL,  18,   0.90,  False),
    ("s4_f16_o32",       S4 | F16 | O32,   N_FULL,  19,   0.90,  False),
    ("s8_b16_o32",       S8 | B16 | O32,   N_FULL,  20,   0.90,  False),
    ("s8_b32_o16",       S8 | B32 | O16,   N_O16,   21,   0.90,  False),
    ("s16_b32_o32",      S16 | B32 | O32,  N_SMALL, 22,   0.90,  False),
    ("f_b32_o32",        F | B32 | O32,    N_SMALL, 23,   0.90,  False),
    ("f_b16_o32",        F | B16 | O32,    N_SMALL, 24,   0.90,  False),
    ("f_f16_o16",        F | F16 | O16,    N_SMALL, 25,   0.90,  False),
    # tie stress: all-zero channel words -> every compare is a tie (soft) / maximally tied (hard)
    ("tie_h_b32",        H | B32 | O32,    N_SMALL, 1,    0.0,   True),
    ("tie_s4_b32",       S4 | B32 | O32,   N_SMALL, 1,    0.0,   True),
    ("tie_s4_b16",       S4 | B16 | O32,   N_SMALL, 1,    0.0,   True),
    ("tie_s4_f16",       S4 | F16 | O32,   N_SMALL, 1,    0.0,   True),
    ("tie_s8_b16_o16",   S8 | B16 | O16,   N_SMALL, 1,    0.0,   True),
    ("tie_f_f16",        F | F16 | O32,    N_SMALL, 1,    0.0,   True),
    # DPX flag: same core as REG in the reference (viterbi.cu:181,192,204)
    ("s4_b16_o32_dpx",   S4 | B16 | O32 | 0x1000, N_SMALL, 26, 0.90, False),
]

# The reference's DPX code paths made live (oracle/_ref/libvitref_dpx.so): tests/golden/ref_vectors_dpx.npz
DPX_CASES = [
    ("dpx_s4_b32",       S4 | B32 | O32,   N_SMALL, 31,   0.90,  False),
    ("dpx_h_b32_o16",    H | B32 | O16,    N_SMALL, 32,   0.80,  False),
    ("dpx_s8_b16",       S8 | B16 | O32,   N_SMALL, 33,   0.90,  False),
    ("dpx_tie_s4_b32",   S4 | B32 | O32,   N_SMALL, 1,    0.0,   True),     # differs from the REG code at phase 0
    ("dpx_tie_s8_b32",   S8 | B32 | O16,   N_SMALL, 1,    0.0,   True),
    ("dpx_tie_s4_b16",   S4 | B16 | O32,   N_SMALL, 1,    0.0,   True),     # int16x2: same table as REG
    ("dpx_tie_f_b32",    F | B32 | O32,    N_SMALL, 1,    0.0,   True),
]
