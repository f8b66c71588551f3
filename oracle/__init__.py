"""CPU oracle (test infrastructure only). See stainx_oracle.c and oracle.py."""
