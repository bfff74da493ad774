"""CPU oracle for the DAMSM hot path -- test infrastructure, never imported by the product path."""
