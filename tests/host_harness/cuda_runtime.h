/* empty stand-in: the host harness compiles the single-thread glue headers with g++ */
