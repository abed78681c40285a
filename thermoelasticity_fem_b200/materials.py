"""Material law of the drop-in API (reference ``materials.py:4-33``)."""
import numpy as np


class LinearThermoElastic:
    """Isotropic linear thermoelastic material; same constructor and attributes as the reference
    (``materials.py:5-16``): ``rho, Y, nu, k, c, alpha, T0`` plus derived ``beta`` (``:18-21``) and
    the 6x6 Mandel-notation elasticity matrix ``mat_C`` (``:23-29``)."""

    def __init__(self, rho, Y, nu, k, c, alpha, T0):
        self.rho = rho
        self.Y = Y
        self.nu = nu
        self.k = k
        self.c = c
        self.alpha = alpha
        self.T0 = T0
        self.beta = None
        self.compute_beta()
        self.mat_C = None
        self.compute_mat_C()

    def lame(self):
        lam = (self.Y * self.nu) / ((1 + self.nu) * (1 - 2 * self.nu))
        mu = self.Y / (2 * (1 + self.nu))
        return lam, mu

    def compute_beta(self):
        lam, mu = self.lame()
        self.beta = (3 * lam + 2 * mu) * self.alpha

    def compute_mat_C(self):
        lam, mu = self.lame()
        self.mat_C = 2 * mu * np.eye(6)
        self.mat_C[:3, :3] += lam

    def table_row(self):
        """Packed row of the device material table: (rho, lambda, mu, k, c, beta, T0, 0)."""
        lam, mu = self.lame()
        return np.array([self.rho, lam, mu, self.k, self.c, self.beta, self.T0, 0.0], dtype=np.float64)


# based on glass SG773 (reference materials.py:33)
glass_SG773 = LinearThermoElastic(rho=2300, Y=64e9, nu=0.1, k=1., c=750., alpha=4e-6, T0=20.)
