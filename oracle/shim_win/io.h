/* empty stand-in for the Windows header the reference generator includes (TEST INFRASTRUCTURE) */
