"""GPU mirrors of the reference's stand-alone Distributed_opt scripts."""
