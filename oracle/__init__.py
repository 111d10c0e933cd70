"""CPU oracle for the RecConv hot path.  TEST INFRASTRUCTURE — never imported by recnext_b200."""
