"""`model` package shadowing the hot-path modules of /root/reference/v2/model."""
