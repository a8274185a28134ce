NAME knap
ROWS
 N obj
 L cap
 G cover
COLUMNS
 M1 'MARKER' 'INTORG'
 a obj -3 cap 4
 a cover 1
 b obj -5 cap 7
 c obj -4 cap 5
 c cover 1
 M2 'MARKER' 'INTEND'
 s cap 1
RHS
 r cap 12
 r cover 1
BOUNDS
 UP b a 3
 UP b b 1
 BV b c
 UP b s 2.5
ENDATA
