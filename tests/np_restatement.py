"""Second, independent restatement of src/engine.rs:572-588 in numpy float32 scalars.

Written without looking at oracle/pbx_oracle.c's structure: every operation is an explicit
np.float32 op in a Python loop, so there is no compiler between the text of the reference
and the arithmetic.  Used only to cross-check the C oracle on small inputs.
"""
import numpy as np

F = np.float32


def u8_to_float(u8s):
    # |v| ((*v as f32 / 255.0) * 2.0) - 1.0          engine.rs:576
    return [((F(int(v)) / F(255.0)) * F(2.0)) - F(1.0) for v in u8s]


def cosine_distance(hash_a, hash_b):
    a = u8_to_float(hash_a)
    b = u8_to_float(hash_b)

    def fold_mag(xs):                         # engine.rs:580
        init = F(0.0)
        for x in xs:
            init = init + x * x
        return init

    magnitude = np.sqrt(fold_mag(a)) * np.sqrt(fold_mag(b))   # engine.rs:581
    if magnitude < F(1e-6):                  # engine.rs:582
        return F(0.0)
    dot = F(0.0)
    for x, y in zip(a, b):                   # engine.rs:585
        dot = dot + (x * y)
    cosine_similarity = dot / magnitude      # engine.rs:586
    m = cosine_similarity if cosine_similarity > F(1e-6) else F(1e-6)
    if np.isnan(cosine_similarity):
        m = F(1e-6)
    return (F(1.0) / m) - F(1.0)             # engine.rs:587


def byte_distance(hash_a, hash_b):
    # fold(0f32, |init, (&a, &b)| init + (a as f32 - b as f32).abs()) / (255f32 * len as f32)      engine.rs:590-592
    init = F(0.0)
    for a, b in zip(hash_a, hash_b):
        init = init + abs(F(int(a)) - F(int(b)))
    return init / (F(255.0) * F(len(hash_a)))


def hamming_distance(hash_a, hash_b):
    # per byte: bits of a ^ b counted into a u8; .sum::<u8>() (wrapping in a release build) / (8f32 * len as f32)   :594-604
    total = 0
    for a, b in zip(hash_a, hash_b):
        diff = int(a) ^ int(b)
        bits_set = 0
        while diff != 0:
            bits_set += diff & 1
            diff >>= 1
        total = (total + bits_set) & 0xFF
    return F(total) / (F(8.0) * F(len(hash_a)))
