from ._preprocessing import normalize, reduce_dim, start_omega
