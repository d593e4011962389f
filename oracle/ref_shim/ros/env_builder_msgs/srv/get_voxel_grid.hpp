// TEST INFRASTRUCTURE - stand-in (unused by the code under test)
