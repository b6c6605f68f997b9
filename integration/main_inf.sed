# SURVEY.md trap T1 / row N3: `--max-chain-skip=infinity` (the spelling of the reference's README, README.md:85-96) is parsed with atoi
# and becomes 0, which on the CPU path disables chaining across any skipped seed.  One line of main.c (main.c:210): "inf" / "infinity"
# -> INT32_MAX, numbers as before.  Applied by oracle/Makefile to the scratch copy the integration binary is built from.
s|opt.max_chain_skip = atoi(o.arg); // --max-chain-skip|opt.max_chain_skip = (o.arg[0] == 'i' \|\| o.arg[0] == 'I') ? 2147483647 : atoi(o.arg); // --max-chain-skip ("inf", "infinity": INT32_MAX)|
