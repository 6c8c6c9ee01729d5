"""
Identities - target identity per read, same surface as /root/reference/badread/identities.py:22-103
(`Identities(mean, stdev, max_identity, output)`, `get_identity()`, `type`, `beta_a`, `beta_b`).
`get_identity(rng)` optionally takes a numpy RandomState so the driver can give every read its own stream;
without it the global numpy stream is used, as in the reference.
"""
import sys

import numpy as np

from .misc import float_to_str, print_in_two_columns


class Identities(object):

    def __init__(self, mean, stdev, max_identity, output=sys.stderr):
        self.mean, self.stdev, self.max_identity = None, None, None
        self.beta_a, self.beta_b = None, None
        print('', file=output)
        if max_identity is None:
            self.type = 'normal'
            self.set_up_normal(mean, stdev, output)
        else:
            self.type = 'beta'
            self.set_up_beta(mean, stdev, max_identity, output)

    def set_up_beta(self, mean, stdev, max_identity, output):
        self.mean = mean / 100.0
        self.stdev = stdev / 100.0
        self.max_identity = max_identity / 100.0
        if self.mean == self.max_identity:
            print(f'Using a constant read identity of {self.mean * 100}%', file=output)
        elif self.stdev == 0.0:
            self.max_identity = self.mean
            print(f'Using a constant read identity of {self.mean * 100}%', file=output)
        else:
            print('Generating read identities from a beta distribution:', file=output)
            self.beta_a, self.beta_b = beta_parameters(mean, stdev, max_identity)
            print_in_two_columns(f'  mean  = {float_to_str(self.mean * 100):>3}%',
                                 f'  max   = {float_to_str(self.max_identity * 100):>3}%',
                                 f'  stdev = {float_to_str(self.stdev * 100):>3}%',
                                 'shape parameters:',
                                 f'  alpha = {self.beta_a:.4e}',
                                 f'  beta  = {self.beta_b:.4e}',
                                 output=output)

    def set_up_normal(self, mean, stdev, output):
        self.mean = mean
        self.stdev = stdev
        if self.stdev == 0.0:
            self.max_identity = self.mean
            print(f'Using a constant read qscore of {self.mean}', file=output)
        else:
            print('Generating read qscores from a normal distribution:', file=output)
            print(f'  mean  = {float_to_str(self.mean):>3}', file=output)
            print(f'  stdev = {float_to_str(self.stdev):>3}', file=output)

    def get_identity(self, rng=None):
        rng = np.random if rng is None else rng
        while True:
            if self.type == 'beta':
                identity = self.get_beta_identity(rng)
            else:
                identity = self.get_normal_identity(rng)
            if 0 <= identity <= 100:
                return identity

    def get_beta_identity(self, rng=np.random):
        if self.mean == self.max_identity:
            return self.mean
        return self.max_identity * rng.beta(self.beta_a, self.beta_b)

    def get_normal_identity(self, rng=np.random):
        qscore = rng.normal(self.mean, self.stdev)
        return 1.0 - 10 ** (-qscore / 10)


def beta_parameters(beta_mean, beta_stdev, beta_max):
    u, s, m = beta_mean, beta_stdev, beta_max
    beta_a = (((1 - (u / m)) / ((s / m) ** 2)) - (m / u)) * ((u / m) ** 2)
    beta_b = beta_a * ((m / u) - 1)
    if beta_a < 0.0 or beta_b < 0.0:
        sys.exit('Error: invalid beta parameters for identity distribution - trying increasing '
                 'the maximum identity or reducing the standard deviation')
    return beta_a, beta_b
