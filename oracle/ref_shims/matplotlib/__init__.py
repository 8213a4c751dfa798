"""Import shim (test infrastructure): the reference imports matplotlib at module scope but the
hot path never plots.  Only used when /root/reference is imported to generate golden fixtures."""
