"""CPU oracle of the kontiki hot path -- TEST INFRASTRUCTURE ONLY (see oracle/README.md)."""
