"""Oracle vs. the independent anchors of SURVEY.md Appendix F and vs. the host tables (CPU)."""
import os

import numpy as np

from conftest import GOLDEN


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def test_schedule_anchor_bits(oracle):
    beta, alpha, acum = oracle.schedule(500)
    # SURVEY.md Appendix F (independent restatement made during the survey)
    assert [hex(v) for v in bits(beta[[0, 1, 2, 249, 499]])] == ["0x38d1b717", "0x3912acb0", "0x393c7dd4", "0x3c24551f", "0x3ca3d70a"]
    assert [hex(v) for v in bits(acum[[0, 1, 249, 498, 499]])] == ["0x3f7ff972", "0x3f7ff047", "0x3e8fb5fd", "0x3bd469eb", "0x3bd02a5c"]
    assert np.array_equal(alpha, (np.float32(1) - beta).astype(np.float32))


def test_schedule_two_restatements_agree(oracle, tabs):
    # oracle: Base twice-precision _linspace restated; host: Float64 formula on Float32 end points
    for T in (5, 500, 1000):
        from igdm_b200 import tables
        b_o, _, a_o = oracle.schedule(T)
        b_h, _, a_h = tables.beta_schedule(T)
        assert np.array_equal(bits(b_o), bits(b_h)), T
        assert np.array_equal(bits(a_o), bits(a_h)), T


def test_embedding_anchor_bits(oracle, tabs):
    pe = oracle.embedding_table(500)
    assert [hex(v) for v in bits(pe[0, :4])] == ["0x3f576aa4", "0x3f0a5140", "0x3f42d671", "0x3f260e19"]
    assert [hex(v) for v in bits(pe[0, 126:128])] == ["0x38e17d42", "0x3f800000"]
    assert [hex(v) for v in bits(pe[249, :4])] == ["0xbf787486", "0x3e76c5a3", "0x3f001662", "0xbf5da6e9"]
    assert [hex(v) for v in bits(pe[499, :4])] == ["0xbeef7fc9", "0xbf6243f2", "0xbf5dcdac", "0x3effa66f"]
    assert np.array_equal(bits(pe), bits(tabs["pe"]))
    # embedding of t=0 is (0,1,0,1,...)
    e0 = oracle.timestep_embedding(0)
    assert np.array_equal(e0[0::2], np.zeros(64, np.float32)) and np.array_equal(e0[1::2], np.ones(64, np.float32))


def test_sampler_scalar_anchor_bits(oracle):
    _, _, acum = oracle.schedule(500)
    assert [hex(v) for v in bits(oracle.sampler_scalars(acum, 500))] == ["0x3f7f2f81", "0x3da33bc7", "0x3da4e408", "0x3f7f2b3e"]
    assert [hex(v) for v in bits(oracle.sampler_scalars(acum, 2))] == ["0x3c7dc584", "0x3f7ff823", "0x3f7ffcb9", "0x3c23da85"]
    tab = oracle.sampler_table(acum)
    # post_var cancels to 1 - alpha_cum[t-1] in exact arithmetic (SURVEY.md trap 4); check in f32 at the anchors
    for t in (2, 500):
        assert tab[t - 1, 3] == np.sqrt(np.float32(1) - acum[t - 2])


def test_apply_noise_closed_form(oracle):
    A, B = oracle.apply_noise_coeffs()
    assert abs(A - 0.079302) < 1e-6 and abs(B - 14.892430) < 1e-6
    img = np.full((64, 64), 0.7)   # the reference's own test input (test/runtests.jl:17)
    eps = np.random.default_rng(7).standard_normal((64, 64))
    out = oracle.apply_noise_f64(img, eps)
    assert np.allclose(out, A * img + B * eps, rtol=0, atol=1e-12)
    assert not np.all(out == img)


def test_philox_known_answer(oracle):
    # Random123 kat_vectors: philox4x32-10, counter = key = 0
    got = oracle.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))[0]
    assert [hex(v) for v in got] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    ones = np.full((1, 4), 0xFFFFFFFF, np.uint32)
    got = oracle.philox4x32_10(ones, np.full((1, 2), 0xFFFFFFFF, np.uint32))[0]
    assert [hex(v) for v in got] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    z = oracle.device_normal(1, np.arange(64), 0)
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02


def test_golden_tables(oracle):
    g = np.load(os.path.join(GOLDEN, "oracle_golden.npz"))
    beta, _, acum = oracle.schedule(500)
    assert np.array_equal(bits(beta), g["beta_bits"]) and np.array_equal(bits(acum), g["acum_bits"])
    assert np.array_equal(bits(oracle.embedding_table(500)[[0, 249, 499]]), g["pe_bits_t1_t250_t500"])
    assert np.array_equal(bits(oracle.sampler_table(acum)), g["samp_bits"])
    assert np.array_equal(oracle.device_normal(3, np.array([0, 1]), 0)[:, :8], g["devnormal_head"])
