"""TEST INFRASTRUCTURE: CPU oracle (plain-C restatement of the reference). Not part of the product."""
