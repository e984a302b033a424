"""CPU oracle of the SSD box codec (test infrastructure, see ssd_codec_oracle.py)."""
