/* oracle/shim -- TEST INFRASTRUCTURE ONLY.  Pf/part.c:11 includes this header
 * but calls nothing from it. */
