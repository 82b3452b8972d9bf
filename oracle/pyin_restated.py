"""CPU oracle for probabilistic YIN (pYIN)  --  TEST INFRASTRUCTURE ONLY.

Restates ``librosa.pyin`` (librosa 0.11; Mauch & Dixon 2014) for the call the reference makes at
``/root/reference/spev_real_metrics.py:369`` (and ``:311`` in the stats pass):

    f0, _, voiced_prob = librosa.pyin(y, fmin=60, fmax=500, sr=22050, hop_length=256)

i.e. frame_length=2048, win_length=1024, n_thresholds=100, beta_parameters=(2, 18),
boltzmann_parameter=2, resolution=0.1, max_transition_rate=35.92, switch_prob=0.01,
no_trough_prob=0.01, fill_na=nan, center=True, pad_mode='constant'.

PARITY PIN STATUS: **unpinned by the reference** (librosa is not vendored and not installable here;
the reference ships no tests).  No second pYIN implementation exists in this container either, so the
restatement is pinned by (tests/test_oracle_pyin.py):
  * the YIN difference function against its O(W*tau) definition (independent brute force);
  * Viterbi against exhaustive path enumeration on small random HMMs;
  * structural properties of the transition matrices (row-stochastic, band-limited);
  * PHYSICAL known answers: harmonic tones / glides of known f0 must decode to that f0 (within one
    10-cent bin) with voiced=True, white noise and silence must decode as unvoiced.
Only scipy's own ``stats.beta`` / ``stats.boltzmann`` / ``signal.get_window`` are reused.
"""
from __future__ import annotations

import numpy as np
import scipy.signal
import scipy.stats

TINY64 = np.finfo(np.float64).tiny


# ---------------------------------------------------------------------------------------------
# YIN front end  (librosa.core.pitch._cumulative_mean_normalized_difference / _parabolic_interpolation)
# ---------------------------------------------------------------------------------------------
def frame_signal(y, frame_length=2048, hop_length=256, center=True):
    y = np.asarray(y)
    if center:
        y = np.pad(y, frame_length // 2, mode="constant")
    n_frames = 1 + (len(y) - frame_length) // hop_length
    idx = (np.arange(n_frames) * hop_length)[None, :] + np.arange(frame_length)[:, None]
    return y[idx]                                              # [frame_length, T]


def cmnd(y_frames, frame_length, win_length, min_period, max_period):
    """Cumulative mean normalised difference, FFT-based exactly like librosa (incl. the
    ``|.| < 1e-6 -> 0`` clean-ups).  y_frames ``[frame_length, T]`` -> ``[max_period-min_period+1, T]``.
    Note the dtype: for float32 audio the FFTs and energies are float32 (numpy >= 2), but the cumulative mean
    divides by an int64 lag vector, so the returned curve -- and everything downstream -- is float64."""
    a = np.fft.rfft(y_frames, frame_length, axis=-2)
    b = np.fft.rfft(y_frames[..., win_length:0:-1, :], frame_length, axis=-2)
    acf_frames = np.fft.irfft(a * b, frame_length, axis=-2)[..., win_length:, :]
    acf_frames[np.abs(acf_frames) < 1e-6] = 0
    energy_frames = np.cumsum(y_frames ** 2, axis=-2)
    energy_frames = energy_frames[..., win_length:, :] - energy_frames[..., :-win_length, :]
    energy_frames[np.abs(energy_frames) < 1e-6] = 0
    yin_frames = energy_frames[..., :1, :] + energy_frames - 2 * acf_frames
    yin_numerator = yin_frames[..., min_period: max_period + 1, :]
    tau_range = np.arange(1, max_period + 1)[:, None]
    cumulative_mean = np.cumsum(yin_frames[..., 1: max_period + 1, :], axis=-2) / tau_range
    yin_denominator = cumulative_mean[..., min_period - 1: max_period, :]
    return yin_numerator / (yin_denominator + np.finfo(yin_denominator.dtype).tiny)


def difference_function_bruteforce(frame, win_length, max_period):
    """d(tau) = sum_{j=1..W} (x_j - x_{j+tau})^2 straight from the definition (float64).  The window
    starts at sample 1: librosa correlates with ``y_frames[win_length:0:-1]`` and differences the
    cumulative energy at ``[tau+W] - [tau]``, i.e. both terms cover samples 1..W (+tau)."""
    x = np.asarray(frame, dtype=np.float64)
    return np.array([np.sum((x[1: win_length + 1] - x[1 + tau: 1 + tau + win_length]) ** 2) for tau in range(max_period + 1)])


def parabolic_interpolation(x):
    """librosa.core.pitch._parabolic_interpolation along axis -2: vertex offset of the parabola through
    (x[i-1], x[i], x[i+1]); 0 at the ends and where |shift| > 1 / flat."""
    x = np.asarray(x)
    shifts = np.zeros_like(x)
    a = x[2:] + x[:-2] - 2 * x[1:-1]
    b = (x[2:] - x[:-2]) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        s = -b / a
    s[np.abs(b) >= np.abs(a)] = 0
    s[~np.isfinite(s)] = 0
    shifts[1:-1] = s
    return shifts


# ---------------------------------------------------------------------------------------------
# HMM pieces  (librosa.sequence.transition_local / transition_loop / viterbi)
# ---------------------------------------------------------------------------------------------
def transition_loop(n_states, prob):
    t = np.empty((n_states, n_states), dtype=np.float64)
    p = np.full(n_states, prob, dtype=np.float64)
    for i in range(n_states):
        t[i] = (1.0 - p[i]) / (n_states - 1)
        t[i, i] = p[i]
    return t


def transition_local(n_states, width, window="triangle"):
    t = np.zeros((n_states, n_states), dtype=np.float64)
    win = scipy.signal.get_window(window, width, fftbins=False)
    lpad = (n_states - width) // 2
    base = np.pad(win, (lpad, n_states - width - lpad))       # util.pad_center
    for i in range(n_states):
        row = np.roll(base, n_states // 2 + i + 1)
        row[min(n_states, i + width // 2 + 1):] = 0           # wrap=False: knock out the wrapped band
        row[: max(0, i - width // 2)] = 0
        t[i] = row
    t /= t.sum(axis=1, keepdims=True)
    return t


def viterbi(prob, transition, p_init):
    """librosa.sequence.viterbi: log-domain with +tiny, argmax ties -> first index.
    prob ``[n_states, T]`` -> state path ``[T]``."""
    return viterbi_log(np.log(prob.T + TINY64), np.log(transition + TINY64), np.log(p_init + TINY64))


def viterbi_log(log_prob, log_trans, log_p_init):
    """The recursion of librosa.sequence._viterbi on ready-made logs: log_prob ``[T, n_states]``,
    log_trans ``[from, to]``."""
    n_steps, n_states = log_prob.shape
    value = np.zeros((n_steps, n_states))
    ptr = np.zeros((n_steps, n_states), dtype=np.int64)
    value[0] = log_prob[0] + log_p_init
    lt = np.ascontiguousarray(log_trans.T)                    # [to, from]
    for t in range(1, n_steps):
        trans_out = value[t - 1][None, :] + lt
        ptr[t] = np.argmax(trans_out, axis=1)
        value[t] = log_prob[t] + trans_out[np.arange(n_states), ptr[t]]
    state = np.zeros(n_steps, dtype=np.int64)
    state[-1] = np.argmax(value[-1])
    for t in range(n_steps - 2, -1, -1):
        state[t] = ptr[t + 1, state[t + 1]]
    return state


def viterbi_bruteforce(prob, transition, p_init):
    """Exhaustive search over all state paths (tiny problems only)."""
    import itertools
    n_states, n_steps = prob.shape
    lt, lp, li = np.log(transition + TINY64), np.log(prob + TINY64), np.log(p_init + TINY64)
    best, arg = -np.inf, None
    for path in itertools.product(range(n_states), repeat=n_steps):
        v = li[path[0]] + lp[path[0], 0]
        for t in range(1, n_steps):
            v += lt[path[t - 1], path[t]] + lp[path[t], t]
        if v > best:
            best, arg = v, path
    return np.array(arg)


# ---------------------------------------------------------------------------------------------
# pYIN
# ---------------------------------------------------------------------------------------------
class PyinConfig:
    def __init__(self, fmin=60.0, fmax=500.0, sr=22050, frame_length=2048, hop_length=256, n_thresholds=100,
                 beta_parameters=(2, 18), boltzmann_parameter=2, resolution=0.1, max_transition_rate=35.92,
                 switch_prob=0.01, no_trough_prob=0.01):
        self.fmin, self.fmax, self.sr = float(fmin), float(fmax), sr
        self.frame_length, self.hop_length = frame_length, hop_length
        self.win_length = frame_length // 2
        self.min_period = int(np.floor(sr / fmax))
        self.max_period = min(int(np.ceil(sr / fmin)), frame_length - self.win_length - 1)
        self.thresholds = np.linspace(0, 1, n_thresholds + 1)
        self.beta_probs = np.diff(scipy.stats.beta.cdf(self.thresholds, beta_parameters[0], beta_parameters[1]))
        self.boltzmann_parameter = boltzmann_parameter
        self.no_trough_prob = no_trough_prob
        self.n_bins_per_semitone = int(np.ceil(1.0 / resolution))
        self.n_pitch_bins = int(np.floor(12 * self.n_bins_per_semitone * np.log2(fmax / fmin))) + 1
        max_semitones_per_frame = round(max_transition_rate * 12 * hop_length / sr)
        self.transition_width = max_semitones_per_frame * self.n_bins_per_semitone + 1
        self.switch_prob = switch_prob
        self.freqs = self.fmin * 2 ** (np.arange(self.n_pitch_bins) / (12 * self.n_bins_per_semitone))

    def transition(self):
        t_local = transition_local(self.n_pitch_bins, self.transition_width)
        t_switch = transition_loop(2, 1 - self.switch_prob)
        return np.kron(t_switch, t_local)

    def p_init(self):
        p = np.zeros(2 * self.n_pitch_bins)
        p[self.n_pitch_bins:] = 1 / self.n_pitch_bins
        return p


def _localmin(x):
    """librosa.util.localmin along axis 0: x[i] < x[i-1] and x[i] <= x[i+1], edges padded."""
    xp = np.pad(x, 1, mode="edge")
    return (x < xp[:-2]) & (x <= xp[2:])


def observation_probs(yin_frames, parabolic_shifts, cfg: PyinConfig):
    """librosa.core.pitch.__pyin_helper.  -> (observation_probs [2*n_bins, T], voiced_prob [T])."""
    n_lags, T = yin_frames.shape
    yin_probs = np.zeros_like(yin_frames)
    thr = cfg.thresholds
    for i in range(T):
        yin_frame = yin_frames[:, i]
        is_trough = _localmin(yin_frame)
        is_trough[0] = yin_frame[0] < yin_frame[1]
        (trough_index,) = np.nonzero(is_trough)
        if len(trough_index) == 0:
            continue
        trough_heights = yin_frame[trough_index]
        trough_thresholds = np.less.outer(trough_heights, thr[1:])
        trough_positions = np.cumsum(trough_thresholds, axis=0) - 1
        n_troughs = np.count_nonzero(trough_thresholds, axis=0)
        trough_prior = scipy.stats.boltzmann.pmf(trough_positions, cfg.boltzmann_parameter, n_troughs)
        trough_prior[~trough_thresholds] = 0
        probs = trough_prior.dot(cfg.beta_probs)
        global_min = np.argmin(trough_heights)
        n_thresholds_below_min = np.count_nonzero(~trough_thresholds[global_min, :])
        probs[global_min] += cfg.no_trough_prob * np.sum(cfg.beta_probs[:n_thresholds_below_min])
        yin_probs[trough_index, i] = probs
    yin_period, frame_index = np.nonzero(yin_probs)
    period_candidates = cfg.min_period + yin_period
    period_candidates = period_candidates + parabolic_shifts[yin_period, frame_index]
    f0_candidates = cfg.sr / period_candidates
    bin_index = 12 * cfg.n_bins_per_semitone * np.log2(f0_candidates / cfg.fmin)
    bin_index = np.clip(np.round(bin_index), 0, cfg.n_pitch_bins).astype(int)
    obs = np.zeros((2 * cfg.n_pitch_bins, T))
    obs[bin_index, frame_index] = yin_probs[yin_period, frame_index]
    voiced_prob = np.clip(np.sum(obs[: cfg.n_pitch_bins, :], axis=0), 0, 1)
    obs[cfg.n_pitch_bins:, :] = (1 - voiced_prob) / cfg.n_pitch_bins
    return obs, voiced_prob


def pyin(y, *, fmin=60.0, fmax=500.0, sr=22050, frame_length=2048, hop_length=256, fill_na=np.nan, return_states=False):
    """-> (f0 [T], voiced_flag [T], voiced_prob [T]) like librosa.pyin."""
    cfg = PyinConfig(fmin=fmin, fmax=fmax, sr=sr, frame_length=frame_length, hop_length=hop_length)
    y_frames = frame_signal(np.asarray(y), frame_length, hop_length)
    yin_frames = cmnd(y_frames, frame_length, cfg.win_length, cfg.min_period, cfg.max_period)
    shifts = parabolic_interpolation(yin_frames)
    obs, voiced_prob = observation_probs(yin_frames, shifts, cfg)
    states = viterbi(obs, cfg.transition(), cfg.p_init())
    f0 = cfg.freqs[states % cfg.n_pitch_bins].copy()
    voiced_flag = states < cfg.n_pitch_bins
    if fill_na is not None:
        f0[~voiced_flag] = fill_na
    if return_states:
        return f0, voiced_flag, voiced_prob, states
    return f0, voiced_flag, voiced_prob
