"""CPU oracle (test infrastructure only — never imported by tfkaldi_b200/)."""
