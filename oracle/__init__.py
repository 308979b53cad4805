"""TEST INFRASTRUCTURE — CPU oracle for the batched UAV stepper.  Never imported by the product."""
