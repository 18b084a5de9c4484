"""Console output of the multi-agent drivers in the reference's line formats (SCvx/utils/multi_agent_logging.py:1-15):
one fixed-width line per ADMM round and a closing summary block.  Only the rendered text is shared with the reference."""

# (label, format spec) of the columns after the iteration counter, in print order
_ROUND_COLUMNS = (("v", "7.3e"), ("slack", "7.3e"), ("p_res", "7.3e"), ("d_res", "7.3e"), ("\u0394x", "6.2e"), ("\u0394s", "6.2e"),
                  ("o", "6.3f"), ("tr", "6.3f"))
_RULE = "=" * 25


def format_iteration(it, nu_norm, slack_norm, primal_res, dual_res, dx, ds, sigma, tr_radius) -> str:
    values = (nu_norm, slack_norm, primal_res, dual_res, dx, ds, sigma, tr_radius)
    cells = [f"Iter {it:2d}"] + [f"{label}={format(val, spec)}" for (label, spec), val in zip(_ROUND_COLUMNS, values)]
    return " | ".join(cells)


def print_iteration(it, nu_norm, slack_norm, primal_res, dual_res, dx, ds, sigma, tr_radius):
    print(format_iteration(it, nu_norm, slack_norm, primal_res, dual_res, dx, ds, sigma, tr_radius))


def print_summary(total_iters, sigma_final, runtime=None):
    lines = ["", "=== SCvx+ADMM Summary ===", f"  Total iterations: {total_iters}", f"  Final time scale o: {sigma_final:.3f}"]
    if runtime is not None:
        lines.append(f"  Total runtime:    {runtime:.2f}s")
    lines += [_RULE, ""]
    print("\n".join(lines))
