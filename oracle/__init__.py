"""Test-infrastructure oracle (see ref_oracle.py header).  Never imported by the product package."""
