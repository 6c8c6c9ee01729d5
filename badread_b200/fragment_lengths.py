"""
FragmentLengths - gamma-distributed fragment lengths, same surface as
/root/reference/badread/fragment_lengths.py:25-64 (`FragmentLengths(mean, stdev, output)`,
`get_fragment_length()`, `gamma_k`, `gamma_t`). The N50 line of the banner needs scipy and is printed when scipy
imports; the ASCII histogram of the reference's banner is presentation only and is not reproduced.

Derived from Badread (Copyright 2018 Ryan Wick, rrwick@gmail.com, https://github.com/rrwick/Badread), which is free
software under the GNU General Public License version 3 or later; this file mirrors the named parts of the
reference's interface and is distributed under the same licence (see LICENSE and NOTICE at the repository root).
"""
import sys

import numpy as np

from .misc import float_to_str, print_in_two_columns


class FragmentLengths(object):

    def __init__(self, mean, stdev, output=sys.stderr):
        self.mean = mean
        self.stdev = stdev
        print('', file=output)
        if self.stdev == 0:
            self.gamma_k, self.gamma_t = None, None
            print(f'Using a constant fragment length of {mean} bp', file=output)
        else:
            print('Generating fragment lengths from a gamma distribution:', file=output)
            gamma_a, gamma_b, self.gamma_k, self.gamma_t = gamma_parameters(mean, stdev)
            n50 = find_n_value(gamma_a, gamma_b, 50)
            n50_text = f'{int(round(n50)):>6}' if n50 is not None else '   n/a'
            print_in_two_columns(f'  mean  = {float_to_str(mean):>6} bp',
                                 f'  stdev = {float_to_str(stdev):>6} bp',
                                 f'  N50   = {n50_text} bp',
                                 'parameters:',
                                 f'  k (shape)     = {self.gamma_k:.4e}',
                                 f'  theta (scale) = {self.gamma_t:.4e}',
                                 output=output)

    def get_fragment_length(self, rng=None):
        rng = np.random if rng is None else rng
        if self.stdev == 0:
            return int(round(self.mean))
        fragment_length = int(round(rng.gamma(self.gamma_k, self.gamma_t)))
        return max(fragment_length, 1)


def gamma_parameters(gamma_mean, gamma_stdev):
    gamma_a = (gamma_mean ** 2) / (gamma_stdev ** 2)
    gamma_b = gamma_mean / (gamma_stdev ** 2)
    gamma_k = (gamma_mean ** 2) / (gamma_stdev ** 2)
    gamma_t = (gamma_stdev ** 2) / gamma_mean
    return gamma_a, gamma_b, gamma_k, gamma_t


def find_n_value(a, b, n):
    """fragment_lengths.py:67-90: bisection on the base-weighted gamma distribution (N50 for the banner)."""
    try:
        import scipy.special
        import scipy.stats
    except ImportError:
        return None

    def integral(x):
        inc = scipy.special.gammaln(a + 1) + np.log(1 - scipy.stats.gamma.cdf(b * x, a + 1))
        return 1.0 - np.exp(inc - scipy.special.gammaln(a + 1))

    target = 1.0 - (n / 100.0)
    bottom, top = 0.0, 1.0
    while integral(top) < target:
        bottom = top
        top *= 2
    guess = (bottom + top) / 2.0
    while True:
        value = integral(guess)
        if top - bottom < 0.01:
            return guess
        if value < target:
            bottom = guess
        else:
            top = guess
        guess = (bottom + top) / 2.0
