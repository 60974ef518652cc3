/* nothing of gsl_randist is used */
