"""Host-side inputs of the low-discrepancy samplers: the Joe-Kuo generator matrices the
reference's Sobol sampler uploads (SobolDetail::SobolMatrices, 256 dimensions x 52 columns; data file
mray_b200/data/sobol_matrices.bin written by oracle/gen_golden_rng.py from the reference's own table) and the
per-generator seeds (consecutive std::mt19937(seed32) draws, Tracer/Random.cu:L917-945)."""
import os
import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
SAMPLER_TYPES = {"Independent": 0, "Sobol": 1, "ZSobol": 2}     # SamplerType (Core/TracerEnums.h)
SOBOL_DIM_COUNT, SOBOL_MATRIX_WIDTH = 256, 52
REFERENCE_SCRAMBLE = 0x100     # MRB_SAMPLER_REFERENCE_SCRAMBLE: the reference's scramble without the final bit reversal


def sampler_code(sampler):
    """'Sobol' / 'ZSobol' / 'Independent', optionally suffixed '+reference' for the reference's exact scramble."""
    if not isinstance(sampler, str):
        return int(sampler)
    name, _, flag = sampler.partition("+")
    return SAMPLER_TYPES[name] | (REFERENCE_SCRAMBLE if flag == "reference" else 0)


def sobol_matrices():
    m = np.fromfile(os.path.join(DATA_DIR, "sobol_matrices.bin"), np.uint32)
    if m.size != SOBOL_DIM_COUNT * SOBOL_MATRIX_WIDTH:
        raise ValueError("sobol_matrices.bin has the wrong size")
    return m


def generator_seeds(seed64, count):
    """LocalState.seed of generator i = i-th draw of std::mt19937(hi32(seed) ^ lo32(seed))."""
    seed32 = ((seed64 >> 32) ^ (seed64 & 0xFFFFFFFF)) & 0xFFFFFFFF
    rs = np.random.RandomState(seed32)          # legacy seeding == init_genrand == std::mt19937(seed32)
    return rs.randint(0, 2 ** 32, size=count, dtype=np.uint64).astype(np.uint32)
