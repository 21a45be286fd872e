// ORACLE — TEST INFRASTRUCTURE ONLY.  (placeholder; the host LBVH rebuild lands with the device builder)
