from ._preprocessing import finish_host_copies, get_state, normalize, reduce_dim, start_omega
