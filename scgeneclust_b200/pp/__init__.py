from ._preprocessing import get_state, normalize, reduce_dim, start_omega
