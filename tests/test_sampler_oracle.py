"""The sampling oracle (oracle_sample) against the reference's own known-answer vectors
(rtp_llm/models_py/bindings/cuda/ops/tests/CudaSamplerTest.cc): top_k == 1 with temperature (:518-568), penalties +
output_all_probs (:905-976), do_sample masks + top-k renormalisation (:986-1057), and the allowed-token sets of the top-k /
top-p tests (:599-657, :659-719). The stochastic draws themselves use flashinfer's Philox stream in the reference and are
not reproducible; everything deterministic around them is pinned here."""
import numpy as np

from oracle import oracle as orc

LOGITS_A = np.array([0, 0, 0, 0.1, 0.2, 0.3, 0, 0, 0, 0.01, 0.987, 0.887, 0.99999, 0.1, 0.2, 0.3, 0, 0, 0.99, 0.989, 0.221, 0, 0, 0.1,
                     0.2, 0.321, 0, 0.4432, 0.44, 0.01, 0.221, 0, 0, 0.1, 0.2, 0.321, 0, 0.4432, 0.44, 0.01], np.float32).reshape(4, 10)


def test_top_k_1_with_temperature_matches_reference_vector():
    tok, _, _, _ = orc.sample(LOGITS_A, top_k=[1, 1, 1, 1], top_p=[1, 1, 1, 1], uniform=[0.3, 0.9, 0.1, 0.7],
                              temperature=[1.0, 10.0, 1.0, 10.0])
    assert tok.tolist() == [5, 2, 7, 7]                     # CudaSamplerTest.cc:564-567


def test_penalties_and_all_probs_match_reference_vector():
    lg = np.array([0.01, 0.88, 0.92, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.01, 0.88, 0.92, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7,
                   0.01, 0.88, 0.92, 0.1, 0.2, 0.3, 0.4, 0.1, 0.1, 0.1, 0.01, 0.88, 0.92, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7],
                  np.float32).reshape(4, 10)
    hist = np.tile(np.array([2, 2, 2, 1, 1, 0], np.int32), (4, 1))
    # sequence_lengths = 5 of a [batch, step + 1 = 6] token buffer: the slot of the token being sampled is not penalised
    # (sampling_penalty_kernels.cu:157-160 skips [input_length, max_input_length))
    tok, tprob, _, probs = orc.sample(lg, top_k=[0] * 4, top_p=[1.0] * 4, uniform=[0.5] * 4, temperature=[1.0] * 4, history=hist,
                                      hist_len=[5] * 4, repetition=[2.4, 1.0, 1.0, 1.2], presence=[0, 0.6, 0, 0.3],
                                      frequency=[0, 0, 0.2, 0.1])
    expect = np.array([0.0693098, 0.0990131, 0.100677, 0.075837, 0.0838128, 0.0926275, 0.102369, 0.113135, 0.125034, 0.138184,
                       0.0703223, 0.0921197, 0.0958792, 0.0769448, 0.0850372, 0.0939806, 0.103865, 0.114788, 0.126861, 0.140203,
                       0.080888, 0.12942, 0.110285, 0.0885056, 0.0978138, 0.108101, 0.11947, 0.0885056, 0.0885056, 0.0885056,
                       0.0715989, 0.0895156, 0.0837425, 0.0783417, 0.0865809, 0.0956867, 0.10575, 0.116872, 0.129164, 0.142748],
                      np.float32).reshape(4, 10)                    # CudaSamplerTest.cc:964-972
    np.testing.assert_allclose(probs, expect, atol=1e-3)
    # cum_log_probs += log p(token): with the reference's sampled tokens (9, 5, 7, 1) the vector at :973-976 follows
    for r, (t, base, want) in enumerate(zip((9, 5, 7, 1), (-1.0, -2.0, -3.0, -3.0), (-2.97917, -4.36467, -5.42469, -5.41334))):
        assert abs(base + np.log(probs[r, t]) - want) < 1e-3
    assert all(0 <= t < 10 for t in tok) and np.allclose(tprob, probs[np.arange(4), tok])


def test_do_sample_mask_and_top_k_renorm_match_reference_vector():
    lg = np.tile(np.array([0.01, 0.8, 0.98, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7], np.float32), (4, 1))
    tok, _, _, probs = orc.sample(lg, top_k=[2] * 4, top_p=[1.0] * 4, uniform=[0.9, 0.1, 0.2, 0.3], temperature=[2.0, 2.0, 4.0, 4.0],
                                  process=[0, 1, 0, 1])           # do_sample = false, true, false, true
    expect = np.zeros((4, 10), np.float32)
    expect[:, 1] = [0.455121, 0.477515, 0.455121, 0.488752]
    expect[:, 2] = [0.544879, 0.522485, 0.544879, 0.511248]       # CudaSamplerTest.cc:1046-1050
    np.testing.assert_allclose(probs, expect, atol=1e-3)
    assert tok.tolist() == [2, 1, 1, 1]                             # u = 0.9 / 0.1 / 0.2 / 0.3 against p(token 1) ~ 0.46-0.49


def test_top_k_and_top_p_keep_exactly_the_reference_allowed_sets():
    # top_k = (1, 1, 3, 2), temperature (1, 10, 1, 10): rows 2 / 3 may only yield {5, 7, 8} / {7, 8} (:655-656)
    for u in np.linspace(0.0, 0.999, 23):
        tok, _, _, probs = orc.sample(LOGITS_A, top_k=[1, 1, 3, 2], top_p=[1.0] * 4, uniform=[u] * 4, temperature=[1.0, 10.0, 1.0, 10.0])
        assert tok[0] == 5 and tok[1] == 2 and tok[2] in (5, 7, 8) and tok[3] in (7, 8)
    assert set(np.flatnonzero(probs[2])) == {5, 7, 8} and set(np.flatnonzero(probs[3])) == {7, 8}
    # top_p keeps the smallest prefix of the sorted probabilities reaching p; u sweeps the whole kept set and nothing else
    lg = np.log(np.array([[0.5, 0.3, 0.1, 0.06, 0.04]], np.float32))
    seen = set()
    for u in np.linspace(0.0, 0.999, 101):
        tok, _, _, probs = orc.sample(lg, top_k=[0], top_p=[0.75], uniform=[u])
        seen.add(int(tok[0]))
    assert seen == {0, 1} and np.allclose(probs[0, :2], [0.625, 0.375], atol=1e-5)
    tok, _, _, probs = orc.sample(lg, top_k=[4], top_p=[0.85], uniform=[0.99])
    assert np.flatnonzero(probs[0]).tolist() == [0, 1, 2] and tok[0] == 2
